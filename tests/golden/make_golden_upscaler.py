"""Golden vectors for the initial-state generator and the IC loss from the reference's OWN `upscaler` classes and
`get_ic_loss` functions (build container only).

    python tests/golden/make_golden_upscaler.py        # rewrites tests/golden/up_*.npz

up_<alias>.npz: a small ragged case -- low-res input, the module's state_dict, `upscaler(low)`, a random upstream
gradient g and the autograd gradients of sum(g * out) for every parameter -- and the script's own `get_ic_loss(model)`
at the size its hard-coded interpolation needs (GS2D:334 (100, 100) from 25^2, GS3D:328 48^3 from 24^3, BUR1:467
(101, 101) from 50^2): the loss, its parameter gradients and the interpolated target.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

CASES = {
    # alias: (small low-res shape, low-res shape get_ic_loss needs, seed)
    "gs2d": ((7, 9), (25, 25), 61),
    "gs3d": ((5, 6, 7), (24, 24, 24), 62),
    "bur1": ((9, 8), (50, 50), 63),
    "bur3": ((6, 11), (50, 50), 64),     # fp64 (BUR3:15)
}


def grads(module):
    out, seen = {}, set()
    for n, p in module.named_parameters():
        if id(p) in seen:
            continue
        seen.add(id(p))
        out[n] = p.grad.detach().numpy().copy()
    return out


def make_case(alias):
    small, ic_shape, seed = CASES[alias]
    mod = mg.load_reference_module(alias)
    dtype = mod._default_dtype
    torch.set_default_dtype(dtype)
    try:
        torch.manual_seed(seed)
        up = mod.upscaler()
    finally:
        torch.set_default_dtype(torch.float32)
    g = torch.Generator().manual_seed(seed)
    low = (torch.rand((1, 2, *small), generator=g, dtype=torch.float64) * 1.6 - 0.8).to(dtype)
    out = up(low)
    gout = (torch.rand(out.shape, generator=g, dtype=torch.float64) - 0.5).to(dtype)
    (out * gout).sum().backward()
    rec = {"low": low.numpy(), "out": out.detach().numpy(), "gout": gout.numpy()}
    for k, v in up.state_dict().items():
        rec["state/" + k] = v.detach().numpy()
    for k, v in grads(up).items():
        rec["grad/" + k] = v
    up.zero_grad()
    ic_low = (torch.rand((1, 2, *ic_shape), generator=g, dtype=torch.float64) * 0.8 + 0.1).to(dtype)
    model = types.SimpleNamespace(UpconvBlock=up, init_state_low=ic_low)
    loss = mod.get_ic_loss(model)
    loss.backward()
    rec["ic_low"] = ic_low.numpy()
    rec["ic_loss"] = np.array(loss.item())
    for k, v in grads(up).items():
        rec["ic_grad/" + k] = v
    # the target the loss compares with, recomputed exactly as the function does (for the drop-in's interpolation)
    import torch.nn.functional as F
    if alias == "gs2d":
        tgt = F.interpolate(ic_low, (100, 100), mode="bicubic")
    elif alias == "gs3d":
        tgt = F.interpolate(ic_low, (48, 48, 48), mode="trilinear")
    else:
        e = torch.cat((ic_low, ic_low[:, :, :, 0:1]), dim=3)
        e = torch.cat((e, e[:, :, 0:1, :]), dim=2)
        tgt = F.interpolate(e, (101, 101), mode="bicubic", align_corners=True)[:, :, :-1, :-1]
    assert abs(float(((up(ic_low) - tgt) ** 2).mean()) - loss.item()) <= 1e-6 * abs(loss.item())
    rec["ic_target"] = tgt.numpy()
    np.savez_compressed(os.path.join(HERE, f"up_{alias}.npz"), **rec)
    print(f"up_{alias}.npz out {tuple(out.shape)} ic_loss {loss.item():.6g} dtype {dtype}")


if __name__ == "__main__":
    torch.set_num_threads(4)
    for alias in CASES:
        make_case(alias)
