"""Fused physics-residual loss: the host side of percnn_phys_loss_fwd / _bwd.

The scripts' `loss_gen(output, loss_generator(dt, dx))` (FWD:288-357 -- the TRAINING loss of the forward-simulation
script, FWD:371-373; GS2D:270-353 and GS3D:286-345, where it is a per-epoch validation metric) pads the whole
trajectory twice, runs two convs, two permute+reshape copies, two Conv1d and ~15 pointwise passes.  Here it is one
kernel for the value (plus the residual gradient) and one for dloss/doutput; like everything in this package it is
CUDA-only and fails loudly otherwise.
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import Tuple

import torch

from . import _lib
from ._lib import check
from .engine import _require_cuda, _stream_ptr

# index of monomial u^a v^b in the cubic packing  c00 c10 c01 c20 c11 c02 c30 c21 c12 c03
_MONO = {(0, 0): 0, (1, 0): 1, (0, 1): 2, (2, 0): 3, (1, 1): 4, (0, 2): 5, (3, 0): 6, (2, 1): 7, (1, 2): 8, (0, 3): 9}


def _cubic(terms) -> Tuple[float, ...]:
    c = [0.0] * 10
    for (a, b), coef in terms.items():
        c[_MONO[(a, b)]] += float(coef)
    return tuple(c)


@dataclasses.dataclass(frozen=True)
class PhysicsSpec:
    """f_q = diff_q * Lap(q) + R_q(u, v) - dq/dt with R_q a bivariate cubic (packing: see _MONO)."""
    diff: Tuple[float, float]
    poly_u: Tuple[float, ...]
    poly_v: Tuple[float, ...]
    dt: float
    dx: float


def lambda_omega_spec(dt: float = 0.0125, dx: float = 0.2) -> PhysicsSpec:
    """FWD:337-340: f_u = 0.1 Lap u + (1-u^2-v^2) u + (u^2+v^2) v - u_t,  f_v = 0.1 Lap v - (u^2+v^2) u + (1-u^2-v^2) v - v_t."""
    ru = _cubic({(1, 0): 1, (3, 0): -1, (1, 2): -1, (2, 1): 1, (0, 3): 1})
    rv = _cubic({(0, 1): 1, (3, 0): -1, (1, 2): -1, (2, 1): -1, (0, 3): -1})
    return PhysicsSpec((0.1, 0.1), ru, rv, float(dt), float(dx))


def gray_scott_spec(Du: float, Dv: float, f: float, k: float, dt: float, dx: float) -> PhysicsSpec:
    """GS2D:321-328 / GS3D:319-326: f_u = Du Lap u - u v^2 + f (1 - u) - u_t,  f_v = Dv Lap v + u v^2 - (f + k) v - v_t."""
    ru = _cubic({(1, 2): -1, (0, 0): f, (1, 0): -f})
    rv = _cubic({(1, 2): 1, (0, 1): -(f + k)})
    return PhysicsSpec((float(Du), float(Dv)), ru, rv, float(dt), float(dx))


_WS = {}


def _workspace(device: torch.device) -> torch.Tensor:
    key = str(device)
    if key not in _WS:
        _WS[key] = torch.empty(int(_lib.lib().percnn_phys_loss_workspace_bytes()), dtype=torch.uint8, device=device)
    return _WS[key]


def _desc(spec: PhysicsSpec, frames: torch.Tensor) -> _lib.PhysLoss:
    _require_cuda(frames, "physics-loss frames")
    if frames.dim() not in (4, 5) or frames.shape[1] != 2:
        raise ValueError(f"expected frames of shape [T, 2, (D,) H, W], got {tuple(frames.shape)}")
    if frames.dtype not in (torch.float32, torch.float64):
        raise TypeError(f"physics loss supports float32/float64, got {frames.dtype}")
    if frames.shape[0] < 3:
        raise ValueError("the physics loss needs at least 3 frames")
    d = _lib.PhysLoss()
    sp = tuple(frames.shape[2:])
    d.ndim = len(sp)
    d.dtype = _lib.F32 if frames.dtype == torch.float32 else _lib.F64
    ext = (1,) + sp if len(sp) == 2 else sp
    for i in range(3):
        d.extent[i] = ext[i]
    d.nframes = frames.shape[0]
    d.device = frames.device.index if frames.device.index is not None else torch.cuda.current_device()
    d.diff[0], d.diff[1] = spec.diff
    for i in range(10):
        d.poly[0][i] = spec.poly_u[i]
        d.poly[1][i] = spec.poly_v[i]
    d.dt, d.dx = spec.dt, spec.dx
    return d


class _PhysicsLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec: PhysicsSpec, frames: torch.Tensor):
        frames = frames.detach().contiguous()
        d = _desc(spec, frames)
        L = _lib.lib()
        need_grad = ctx.needs_input_grad[1]
        resid = torch.empty((frames.shape[0] - 2, *frames.shape[1:]), dtype=frames.dtype, device=frames.device) if need_grad else None
        loss = torch.empty((), dtype=frames.dtype, device=frames.device)
        with torch.cuda.device(frames.device):
            check(L.percnn_phys_loss_fwd(ctypes.byref(d), frames.data_ptr(), None if resid is None else resid.data_ptr(),
                                         loss.data_ptr(), _workspace(frames.device).data_ptr(), _stream_ptr(frames.device)))
        ctx.spec = spec
        if need_grad:
            ctx.save_for_backward(frames, resid)
        return loss

    @staticmethod
    def backward(ctx, g_loss: torch.Tensor):
        frames, resid = ctx.saved_tensors
        d = _desc(ctx.spec, frames)
        g = torch.empty_like(frames)
        gscale = g_loss.detach().to(frames.dtype).reshape(1).contiguous()
        with torch.cuda.device(frames.device):
            check(_lib.lib().percnn_phys_loss_bwd(ctypes.byref(d), frames.data_ptr(), resid.data_ptr(), gscale.data_ptr(),
                                                  g.data_ptr(), _stream_ptr(frames.device)))
        return None, g


def physics_loss(frames: torch.Tensor, spec: PhysicsSpec) -> torch.Tensor:
    """mse(f_u, 0) + mse(f_v, 0) of the un-padded trajectory `frames` [T, 2, (D,) H, W] (consecutive frames one dt
    apart), exactly as `loss_gen(output, loss_func)` computes it, differentiable w.r.t. `frames`."""
    return _PhysicsLoss.apply(spec, frames)


@torch.no_grad()
def physics_residuals(frames: torch.Tensor, spec: PhysicsSpec):
    """(f_u, f_v), each [T-2, 1, (D,) H, W], on the un-padded periodic grid (the reference's `get_phy_Loss` returns
    the same values on extent+1 points per axis, the last being the periodic image of the first)."""
    frames = frames.contiguous()
    d = _desc(spec, frames)
    resid = torch.empty((frames.shape[0] - 2, *frames.shape[1:]), dtype=frames.dtype, device=frames.device)
    loss = torch.empty((), dtype=frames.dtype, device=frames.device)
    with torch.cuda.device(frames.device):
        check(_lib.lib().percnn_phys_loss_fwd(ctypes.byref(d), frames.data_ptr(), resid.data_ptr(), loss.data_ptr(),
                                              _workspace(frames.device).data_ptr(), _stream_ptr(frames.device)))
    # resid = 2 w f / N  ->  f
    n = (frames.shape[0] - 2)
    w = torch.ones(frames.shape[2:], dtype=frames.dtype, device=frames.device)
    for ax, ext in enumerate(frames.shape[2:]):
        n *= ext + 1
        idx = [slice(None)] * w.dim()
        idx[ax] = 0
        w[tuple(idx)] *= 2
    f = resid * (n / 2.0) / w
    return f[:, 0:1], f[:, 1:2]


class LossGenerator(torch.nn.Module):
    """Drop-in for the scripts' `loss_generator(dt, dx)` (FWD:265-286, GS2D:241-262, GS3D:264-283): holds the
    constants; the arithmetic is in the fused kernels."""

    def __init__(self, spec: PhysicsSpec):
        super().__init__()
        self.spec = spec

    def forward(self, output: torch.Tensor) -> torch.Tensor:
        return physics_loss(output, self.spec)

    def get_phy_Loss(self, output):
        """(f_u, f_v) like the reference's method (GS2D:270-329, FWD:288-342, GS3D:286-327): `output` is the trajectory
        PADDED periodically by 2 cells before and 3 after every spatial axis (what `loss_gen` builds, GS2D:343-346);
        each residual field has extent + 1 points per axis, the last being the periodic image of the first.  Computed by
        the fused residual kernel on the un-padded grid (no gradient: the differentiable path is `loss_gen`)."""
        nsp = output.dim() - 2
        core = (slice(None), slice(None)) + (slice(2, -3),) * nsp
        frames = output[core].contiguous()
        for ax in range(nsp):   # the padding must be the periodic one, or the un-padded grid does not describe `output`
            lead = [slice(None)] * output.dim()
            lead[2 + ax] = slice(0, 2)
            src = [slice(None)] * output.dim()
            src[2 + ax] = slice(-5, -3)
            if not torch.allclose(output[tuple(lead)], output[tuple(src)]):
                raise ValueError("get_phy_Loss expects the periodic (2, 3) padding of loss_gen (GS2D:343-346)")
        f_u, f_v = physics_residuals(frames, self.spec)
        for ax in range(nsp):
            first = [slice(None)] * f_u.dim()
            first[2 + ax] = slice(0, 1)
            f_u = torch.cat((f_u, f_u[tuple(first)]), dim=2 + ax)
            f_v = torch.cat((f_v, f_v[tuple(first)]), dim=2 + ax)
        return f_u, f_v
