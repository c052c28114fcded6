"""GPU parity tests of the fused physics-residual loss (percnn_phys_loss_fwd / _bwd).

Against (1) vectors recorded from the reference's own `loss_generator` + `loss_gen` / `loss_func` on the reference's
own trajectories (tests/golden/phys_*.npz: value, dloss/doutput, end-to-end gradients through the rollout), and
(2) the numpy oracle on ragged sizes.  Tolerances: fp64 1e-11; fp32 1e-5 (value) / 2e-5 (gradients) -- the
reference's own fp32-vs-fp64 distance on these cases is 1e-7 / 2e-7.
"""
import os

import numpy as np
import pytest
import torch

from oracle import percnn_oracle as po
from percnn_b200 import _lib, losses
from tests.helpers import GOLDEN, make_cell, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _variant_module(tag):
    from percnn_b200.variants import gs2d, gs3d, lambda_omega_fwd
    return {"fwd": lambda_omega_fwd, "gs2d": gs2d, "gs3d": gs3d}[tag]


def _load(tag):
    z = np.load(os.path.join(GOLDEN, f"phys_{tag}.npz"))
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    grads = {k[len("grad/"):]: z[k] for k in z.files if k.startswith("grad/")}
    return z, params, grads


@pytest.mark.parametrize("tag", ["fwd", "gs2d", "gs3d"])
def test_physics_loss_matches_reference_golden(tag):
    z, _, _ = _load(tag)
    mod = _variant_module(tag)
    fp64 = z["h0"].dtype == np.float64
    out = torch.from_numpy(z["traj"]).to(DEV).requires_grad_(True)
    gen = mod.loss_generator()
    fn = mod.loss_func if tag == "gs3d" else mod.loss_gen
    loss = fn(out, gen)
    assert loss.dim() == 0 and loss.dtype == out.dtype
    assert abs(loss.item() - float(z["loss"])) <= (1e-12 if fp64 else 1e-5) * abs(float(z["loss"]))
    loss.backward()
    assert rel_l2(out.grad.cpu().numpy(), z["g_traj"]) <= (1e-11 if fp64 else 2e-5)
    assert torch.count_nonzero(out.grad[-1]) == 0


@pytest.mark.parametrize("tag", ["fwd", "gs2d", "gs3d"])
def test_training_step_on_physics_loss_matches_reference(tag):
    """FWD:366-373: output = cat(model()); loss = loss_gen(output, loss_func); loss.backward() -- rollout through the
    fused cell, fused loss, dense gradient into the hand-derived adjoint; gradients vs the reference's autograd."""
    z, params, grads = _load(tag)
    mod = _variant_module(tag)
    fp64 = z["h0"].dtype == np.float64
    cell = make_cell(tag)
    cell.load_state_dict(params, strict=True)
    cell = cell.to(DEV)
    h0 = torch.from_numpy(z["h0"]).to(DEV).requires_grad_(True)
    nstep = int(z["nstep"])
    states = cell.rollout(h0, nstep)
    gen = mod.loss_generator()
    loss = (mod.loss_func if tag == "gs3d" else mod.loss_gen)(states, gen)
    (float(z["gscale"]) * loss).backward()
    assert abs(loss.item() - float(z["loss"])) <= (1e-11 if fp64 else 2e-5) * abs(float(z["loss"]))
    gtol = 1e-10 if fp64 else 5e-5
    assert rel_l2(h0.grad.cpu().numpy(), z["g_h0"]) <= gtol
    named = dict(cell.named_parameters())
    for k, ref in grads.items():
        got = named[k].grad
        assert got is not None, k
        tol = gtol if ref.ndim > 0 or fp64 else 5e-4
        assert rel_l2(got.cpu().numpy(), ref) <= tol, (k, rel_l2(got.cpu().numpy(), ref))


@pytest.mark.parametrize("tag,shape,nframes", [("fwd", (17, 23), 4), ("fwd", (5, 5), 3), ("gs2d", (33, 20), 6),
                                                ("gs3d", (5, 6, 7), 4), ("gs3d", (9, 8, 16), 3)])
def test_physics_loss_matches_numpy_oracle_on_ragged_sizes(tag, shape, nframes):
    g = torch.Generator().manual_seed(5)
    dtype = torch.float64 if tag == "fwd" else torch.float32
    out_cpu = (torch.rand((nframes, 2, *shape), generator=g, dtype=torch.float64) - (0.5 if tag == "fwd" else 0.0)).to(dtype)
    want_l, want_g = po.phys_loss_np(out_cpu.numpy(), tag)
    out = out_cpu.to(DEV).requires_grad_(True)
    spec = _variant_module(tag).loss_generator().spec
    loss = losses.physics_loss(out, spec)
    (1.5 * loss).backward()
    fp64 = dtype == torch.float64
    assert abs(loss.item() - want_l) <= (1e-12 if fp64 else 1e-5) * abs(want_l)
    assert rel_l2(out.grad.cpu().numpy(), 1.5 * want_g) <= (1e-11 if fp64 else 2e-5)
    # residuals on the periodic grid reproduce the loss with the double-counting weights
    fu, fv = losses.physics_residuals(out.detach(), spec)
    assert fu.shape == (nframes - 2, 1, *shape)
    w = np.ones(shape)
    n = nframes - 2
    for ax, ext in enumerate(shape):
        n *= ext + 1
        idx = [slice(None)] * len(shape)
        idx[ax] = 0
        w[tuple(idx)] *= 2
    rec = float((w * (fu.cpu().numpy()[:, 0].astype(np.float64) ** 2 + fv.cpu().numpy()[:, 0].astype(np.float64) ** 2)).sum() / n)
    assert abs(rec - want_l) <= (1e-11 if fp64 else 2e-5) * abs(want_l)


def test_physics_loss_no_grad_and_errors():
    spec = losses.lambda_omega_spec()
    out = torch.rand((4, 2, 8, 8), dtype=torch.float64, device=DEV)
    with torch.no_grad():
        l0 = losses.physics_loss(out, spec)
    l1 = losses.physics_loss(out.clone().requires_grad_(True), spec)
    assert l0.item() == l1.item() and not l0.requires_grad and l1.requires_grad
    with pytest.raises(RuntimeError):
        losses.physics_loss(out.cpu(), spec)                      # no CPU fallback
    with pytest.raises(ValueError):
        losses.physics_loss(out[:2], spec)                        # needs >= 3 frames
    with pytest.raises(ValueError):
        losses.physics_loss(out[:, :1], spec)                     # 2 fields
    with pytest.raises(ValueError):                 # get_phy_Loss wants the periodically PADDED trajectory (FWD:344-350)
        from percnn_b200.variants import lambda_omega_fwd
        lambda_omega_fwd.loss_generator().get_phy_Loss(out)
