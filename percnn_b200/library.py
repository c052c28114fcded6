"""Stage-2 library of candidate terms: host side of percnn_library_terms / _theta.

Drop-in for `Loss_generator` of DataDrivenDiscoveryOfPDEs/*/Stage-2/derivatives.py (BUR2d:84-214; the lambda-omega twin
names the method `get_library`) and for the column construction of PDE_FIND_u.py / PDE_FIND_v.py (`gen_library`,
PDE_FIND_u.py:185-193, and the eval'd products PDE_FIND_u.py:246-259).  The sparse regression itself (STRidge, a
`numpy.linalg.lstsq` loop on a 70-column matrix) stays on the host, as in the reference.  CUDA only.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List

import torch

from . import _lib
from ._lib import check
from .engine import _require_cuda, _stream_ptr

TERMS = ("f_u", "f_v", "u", "v", "u_t", "v_t", "u_x", "u_y", "v_x", "v_y", "lap_u", "lap_v")
LIST_A = ["ones", "u", "v", "u**2", "u*v", "v**2", "u**3", "u**2*v", "u*v**2", "v**3"]     # PDE_FIND_u.py:187
LIST_B = ["ones", "u_x", "u_y", "v_x", "v_y", "lap_u", "lap_v"]                             # PDE_FIND_u.py:188


def gen_library() -> List[str]:
    """PDE_FIND_u.py:185-193: the 70 column names, A-major."""
    return [a + "*" + b for a in LIST_A for b in LIST_B]


class Loss_generator(torch.nn.Module):
    """`Loss_generator(dt, dx)` (BUR2d:84-127).  kind: 'burgers' (residual of BUR2d:189-192) or 'lo' (LO2d:188-192)."""

    def __init__(self, dt=0.00025, dx=1.0 / 100, kind: str = "burgers"):
        super().__init__()
        self.dt, self.dx, self.dy = dt, dx, dx
        self.kind = {"burgers": 0, "lo": 1}[kind]

    def _desc(self, frames: torch.Tensor) -> _lib.Library:
        _require_cuda(frames, "library trajectory")
        if frames.dim() != 4 or frames.shape[1] != 2:
            raise ValueError(f"expected a trajectory of shape [T, 2, H, W], got {tuple(frames.shape)}")
        if frames.dtype not in (torch.float32, torch.float64):
            raise TypeError(f"the library supports float32/float64, got {frames.dtype}")
        d = _lib.Library()
        d.dtype = _lib.F32 if frames.dtype == torch.float32 else _lib.F64
        d.kind, d.H, d.W, d.nframes = self.kind, frames.shape[2], frames.shape[3], frames.shape[0]
        d.device = frames.device.index if frames.device.index is not None else torch.cuda.current_device()
        d.dt, d.dx = float(self.dt), float(self.dx)
        return d

    @torch.no_grad()
    def terms(self, output: torch.Tensor) -> torch.Tensor:
        """[12, T-2, 1, H+1, W+1] (order: TERMS) from the UN-padded periodic trajectory `output` [T, 2, H, W]."""
        frames = output.detach().contiguous()
        d = self._desc(frames)
        T, _, H, W = frames.shape
        out = torch.empty((len(TERMS), T - 2, 1, H + 1, W + 1), dtype=frames.dtype, device=frames.device)
        with torch.cuda.device(frames.device):
            check(_lib.lib().percnn_library_terms(ctypes.byref(d), frames.data_ptr(), out.data_ptr(), _stream_ptr(frames.device)))
        return out

    def library_from_periodic(self, output: torch.Tensor) -> Dict[str, torch.Tensor]:
        t = self.terms(output)
        lib = {name: t[i] for i, name in enumerate(TERMS)}
        lib["ones"] = torch.ones_like(lib["u"])
        return lib

    def get_phy_residual(self, output: torch.Tensor) -> Dict[str, torch.Tensor]:
        """BUR2d:129-199: `output` is the trajectory PADDED by (2, 3) periodically, as `get_residual_mse` builds it
        (BUR2d:207-208); returns the dict of [T-2, 1, H+1, W+1] fields."""
        core = output[:, :, 2:-3, 2:-3]
        if not (torch.equal(output[:, :, 0:2, 2:-3], core[:, :, -2:, :]) and torch.equal(output[:, :, 2:-3, 0:2], core[:, :, :, -2:])):
            raise ValueError("get_phy_residual expects the periodic (2, 3) padding of get_residual_mse (BUR2d:207-208)")
        return self.library_from_periodic(core)

    get_library = get_phy_residual     # the lambda-omega script's name for the same method (LO2d:128)

    def get_residual_mse(self, output: torch.Tensor):
        """BUR2d:202-217 on the UN-padded trajectory: (mse(f_u, 0), mse(f_v, 0))."""
        t = self.terms(output)
        return t[0].double().pow(2).mean().to(t.dtype), t[1].double().pow(2).mean().to(t.dtype)

    @torch.no_grad()
    def theta(self, terms: torch.Tensor, idx: torch.Tensor):
        """(lhs [n, 70] fp64, rhs [n, 2] fp64 = (u_t, v_t)) at the flattened sample points `idx` (PDE_FIND_u.py:246-259:
        `terms_dict[key][idx, :]`, the eval'd column products, `rhs = u_t`).  `terms` = self.terms(output)."""
        _require_cuda(terms, "library terms")
        npts = terms[0].numel()
        idx = idx.to(device=terms.device, dtype=torch.int64).contiguous()
        if idx.numel() < 1 or int(idx.min()) < 0 or int(idx.max()) >= npts:
            raise IndexError("sample indices out of range")
        T2, _, H1, W1 = terms.shape[1:]
        d = _lib.Library()
        d.dtype = _lib.F32 if terms.dtype == torch.float32 else _lib.F64
        d.kind, d.H, d.W, d.nframes = self.kind, H1 - 1, W1 - 1, T2 + 2
        d.device = terms.device.index if terms.device.index is not None else torch.cuda.current_device()
        d.dt, d.dx = float(self.dt), float(self.dx)
        n = idx.numel()
        lhs = torch.empty((n, 70), dtype=torch.float64, device=terms.device)
        rhs = torch.empty((n, 2), dtype=torch.float64, device=terms.device)
        with torch.cuda.device(terms.device):
            check(_lib.lib().percnn_library_theta(ctypes.byref(d), terms.contiguous().data_ptr(), idx.data_ptr(), n, lhs.data_ptr(),
                                                  rhs.data_ptr(), _stream_ptr(terms.device)))
        return lhs, rhs
