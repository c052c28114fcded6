"""Fused initial-state generator + IC loss against the reference's own `upscaler` / `get_ic_loss`
(tests/golden/up_*.npz), against the oracle on ragged sizes, and slab mode against the whole-grid call."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import percnn_oracle as po
from percnn_b200 import upscaler as up
from percnn_b200.variants import burgers_stage1, burgers_stage3, gs2d, gs3d
from tests.helpers import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MODS = {"gs2d": gs2d, "gs3d": gs3d, "bur1": burgers_stage1, "bur3": burgers_stage3}
KIND = {"gs2d": "gs2d", "gs3d": "gs3d", "bur1": "stage", "bur3": "stage"}


def _module(alias, z):
    m = MODS[alias].upscaler()
    if z["out"].dtype == np.float64:
        m = m.double()
    sd = {k[len("state/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state/")}
    m.load_state_dict(sd, strict=True)
    return m.to(DEV), sd


@pytest.mark.parametrize("alias", ["gs2d", "gs3d", "bur1", "bur3"])
def test_fused_upscaler_matches_the_reference_module(alias):
    z = np.load(os.path.join(GOLDEN, f"up_{alias}.npz"))
    f64 = z["out"].dtype == np.float64
    m, _ = _module(alias, z)
    out = m(torch.from_numpy(z["low"]).to(DEV))
    assert tuple(out.shape) == z["out"].shape
    assert rel_l2(out.detach().cpu().numpy(), z["out"]) <= (1e-12 if f64 else 2e-6)
    (out * torch.from_numpy(z["gout"]).to(DEV)).sum().backward()
    named = dict(m.named_parameters())
    keys = [k for k in z.files if k.startswith("grad/")]
    assert keys
    for k in keys:
        name = k[len("grad/"):]
        assert rel_l2(named[name].grad.cpu().numpy(), z[k]) <= (1e-11 if f64 else 2e-5), name


@pytest.mark.parametrize("alias", ["gs2d", "gs3d", "bur1", "bur3"])
def test_fused_ic_loss_matches_get_ic_loss(alias):
    z = np.load(os.path.join(GOLDEN, f"up_{alias}.npz"))
    f64 = z["out"].dtype == np.float64
    m, _ = _module(alias, z)
    low = torch.from_numpy(z["ic_low"]).to(DEV)
    model = types.SimpleNamespace(UpconvBlock=m, init_state_low=low)
    mode = {"gs2d": "bicubic", "gs3d": "trilinear"}.get(alias, "bicubic_periodic")
    tgt = up.ic_target(low, mode, z["ic_target"].shape[2:])
    assert rel_l2(tgt.cpu().numpy(), z["ic_target"]) <= (1e-12 if f64 else 2e-6)
    loss = MODS[alias].get_ic_loss(model)
    assert abs(loss.item() - float(z["ic_loss"])) <= (1e-11 if f64 else 1e-5) * abs(float(z["ic_loss"]))
    (0.25 * loss).backward()          # GS2D:406 weights the IC loss by 0.25: the upstream scalar reaches the kernels
    named = dict(m.named_parameters())
    for k in [k for k in z.files if k.startswith("ic_grad/")]:
        name = k[len("ic_grad/"):]
        assert rel_l2(named[name].grad.cpu().numpy(), 0.25 * z[k]) <= (1e-10 if f64 else 3e-5), name


@pytest.mark.parametrize("alias,low_shape,dtype", [
    ("gs2d", (1, 1), torch.float32), ("gs2d", (3, 2), torch.float64), ("gs2d", (13, 31), torch.float32),
    ("gs3d", (1, 1, 1), torch.float32), ("gs3d", (2, 3, 5), torch.float64), ("gs3d", (7, 9, 6), torch.float32),
    ("bur1", (1, 3), torch.float32), ("bur1", (17, 5), torch.float64), ("bur1", (25, 26), torch.float32),
])
def test_fused_upscaler_matches_the_oracle_on_ragged_sizes(alias, low_shape, dtype):
    torch.manual_seed(hash((alias, low_shape)) % 1000)
    m = MODS[alias].upscaler().to(dtype)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(3.0)
    sd = {k: v.detach().clone().double() for k, v in m.state_dict().items()}
    low = torch.rand((1, 2, *low_shape), dtype=torch.float64) * 2 - 1
    prm = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = po.upscaler_torch(low, prm, KIND[alias])
    gout = torch.rand(want.shape, dtype=torch.float64) - 0.5
    (want * gout).sum().backward()
    m = m.to(DEV)
    out = m(low.to(DEV, dtype))
    tol, gtol = (1e-12, 1e-10) if dtype == torch.float64 else (2e-6, 3e-5)
    assert rel_l2(out.detach().cpu().double().numpy(), want.detach().numpy()) <= tol
    (out * gout.to(DEV, dtype)).sum().backward()
    seen = set()
    for name, p in m.named_parameters():
        if id(p) in seen:
            continue
        seen.add(id(p))
        ref = prm[name].grad.numpy()
        got = p.grad.cpu().double().numpy()
        assert np.abs(got - ref).max() <= gtol * max(np.abs(ref).max(), 1e-3), name


def test_upscaler_gradients_are_deterministic():
    torch.manual_seed(3)
    m = gs3d.upscaler().to(DEV)
    low = torch.rand((1, 2, 6, 7, 8), device=DEV)
    g = torch.rand((1, 2, 12, 14, 16), device=DEV)
    res = []
    for _ in range(3):
        m.zero_grad()
        (m(low) * g).sum().backward()
        res.append(torch.cat([p.grad.reshape(-1) for p in m.up_parameters()]).clone())
    assert torch.equal(res[0], res[1]) and torch.equal(res[0], res[2])


@pytest.mark.parametrize("nslab", [2, 3])
def test_slab_mode_equals_the_whole_grid(nslab):
    """Every rank generates its own planes of h0 from the replicated low-res input and owns the parameter sums of its
    planes: slabs tile the whole-grid result bit for bit, partial gradients add up, planes outside the global grid
    are never read (they are NaN here)."""
    torch.manual_seed(5)
    m = gs3d.upscaler().to(DEV)
    Dl, Hl, Wl = 6, 5, 8
    low = torch.rand((1, 2, Dl, Hl, Wl), device=DEV)
    flat = up._pack(m.up_parameters(), torch.float32)
    geo = m.geometry((Dl, Hl, Wl), torch.float32, torch.device(DEV))
    whole, mid = up.upscaler_fwd(geo, flat, low)
    D, H, W = geo.out_shape
    g = torch.rand((2, D, H, W), device=DEV) - 0.5
    gp_whole = up.upscaler_bwd(geo, flat, low, mid, g.contiguous().data_ptr())
    nz = D // nslab
    total = torch.zeros_like(gp_whole, dtype=torch.float64)
    for r in range(nslab):
        z0 = r * nz
        buf = torch.full((2, nz + 4, H, W), float("nan"), device=DEV)      # the slab layout: 2 ghost planes per side
        sgeo = m.geometry((Dl, Hl, Wl), torch.float32, torch.device(DEV), out_z0=z0, out_nz=nz, out_field_stride=(nz + 4) * H * W)
        _, smid = up.upscaler_fwd(sgeo, flat, low, out=buf[:, 2:])
        assert torch.equal(buf[:, 2:nz + 2], whole[0][:, z0:z0 + nz])
        gb = torch.full((2, nz + 4, H, W), float("nan"), device=DEV)
        lo, hi = max(z0 - 2, 0), min(z0 + nz + 2, D)
        gb[:, 2 - (z0 - lo):2 + nz + (hi - z0 - nz)] = g[:, lo:hi]
        part = up.upscaler_bwd(sgeo, flat, low, smid, gb[:, 2:].data_ptr())
        assert torch.isfinite(part).all()
        total += part.double()
    assert rel_l2(total.cpu().numpy(), gp_whole.double().cpu().numpy()) <= 1e-6
