"""Fused initial-state generator and IC loss: host side of percnn_upscaler_* / percnn_mse_*.

The scripts build the full-resolution initial state with `self.UpconvBlock(self.init_state_low)` (GS2D:164, GS3D:186,
BUR1:277) and regularise it with `get_ic_loss(model)` (GS2D:331-338).  `FusedUpscaler` keeps the reference's
sub-modules purely as parameter containers (same `state_dict` keys, same initialisation) and runs the arithmetic in
the library: two kernels forward (the second transposed conv and the 1x1 conv are folded into one), and a hand-derived
adjoint whose parameter sums are deterministic.  CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import check
from .engine import _require_cuda, _stream_ptr

_WS = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = str(device)
    if key not in _WS or _WS[key].numel() < nbytes:
        _WS[key] = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
    return _WS[key]


def _device_index(device: torch.device) -> int:
    return device.index if device.index is not None else torch.cuda.current_device()


class UpscalerGeometry:
    """percnn_upscaler_t plus the sizes the library derives from it."""

    def __init__(self, ndim: int, channels: int, act: str, layers: int, stride2: int, low_shape: Sequence[int], dtype: torch.dtype,
                 device: torch.device, out_z0: int = 0, out_nz: int = 0, out_field_stride: int = 0):
        if dtype not in (torch.float32, torch.float64):
            raise TypeError(f"upscaler supports float32/float64, got {dtype}")
        if len(low_shape) != ndim:
            raise ValueError(f"expected {ndim} spatial dims, got {tuple(low_shape)}")
        d = _lib.Upscaler()
        d.ndim, d.dtype, d.channels = ndim, (_lib.F32 if dtype == torch.float32 else _lib.F64), channels
        d.act = {"sigmoid": 0, "tanh": 1}[act]
        d.layers, d.stride2, d.device = layers, stride2, _device_index(device)
        ext = (1,) + tuple(low_shape) if ndim == 2 else tuple(low_shape)
        for i in range(3):
            d.low_extent[i] = int(ext[i])
        d.out_z0, d.out_nz, d.out_field_stride = int(out_z0), int(out_nz), int(out_field_stride)
        npar, mid, ws = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_size_t()
        out = (ctypes.c_int64 * 3)()
        check(_lib.lib().percnn_upscaler_sizes(ctypes.byref(d), ctypes.byref(npar), ctypes.byref(mid), out, ctypes.byref(ws)))
        self.desc = d
        self.nparams, self.mid_elems, self.ws_bytes = int(npar.value), int(mid.value), int(ws.value)
        self.out_shape = tuple(int(v) for v in out)[3 - ndim:]
        self.dtype, self.device = dtype, device


def _pack(params: Sequence[torch.Tensor], dtype: torch.dtype) -> torch.Tensor:
    return torch.cat([p.detach().reshape(-1).to(dtype) for p in params]).contiguous()


def upscaler_fwd(geo: UpscalerGeometry, flat: torch.Tensor, low: torch.Tensor, out: Optional[torch.Tensor] = None):
    """h0 (or this rank's planes of it, written into `out` when given) and the activation tape `mid`."""
    _require_cuda(low, "init_state_low")
    low = low.detach().to(geo.dtype).contiguous()
    if flat.numel() != geo.nparams:
        raise ValueError(f"upscaler parameter packing has {flat.numel()} values, expected {geo.nparams}")
    mid = torch.empty(geo.mid_elems, dtype=geo.dtype, device=low.device)
    if out is None:
        out = torch.empty((1, 2, *geo.out_shape), dtype=geo.dtype, device=low.device)
    with torch.cuda.device(low.device):
        check(_lib.lib().percnn_upscaler_fwd(ctypes.byref(geo.desc), flat.data_ptr(), low.data_ptr(), mid.data_ptr(), out.data_ptr(),
                                             _workspace(low.device, geo.ws_bytes).data_ptr(), _stream_ptr(low.device)))
    return out, mid


def upscaler_bwd(geo: UpscalerGeometry, flat: torch.Tensor, low: torch.Tensor, mid: torch.Tensor, g_ptr: int,
                 g_params: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Flat parameter gradient for the upstream gradient at device address `g_ptr` (layout: see the header)."""
    accumulate = g_params is not None
    if g_params is None:
        g_params = torch.empty(geo.nparams, dtype=geo.dtype, device=low.device)
    with torch.cuda.device(low.device):
        check(_lib.lib().percnn_upscaler_bwd(ctypes.byref(geo.desc), flat.data_ptr(), low.data_ptr(), mid.data_ptr(), g_ptr,
                                             g_params.data_ptr(), int(accumulate), _workspace(low.device, geo.ws_bytes).data_ptr(),
                                             _stream_ptr(low.device)))
    return g_params


class _UpscalerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: "FusedUpscaler", low: torch.Tensor, *params: torch.Tensor):
        if low.dim() != 2 + module.up_ndim or low.shape[0] != 1 or low.shape[1] != 2:
            raise ValueError(f"expected init_state_low of shape [1, 2, ...{module.up_ndim} dims], got {tuple(low.shape)}")
        dtype = params[0].dtype
        geo = module.geometry(tuple(low.shape[2:]), dtype, low.device)
        flat = _pack(params, dtype)
        out, mid = upscaler_fwd(geo, flat, low)
        ctx.geo, ctx.shapes = geo, [p.shape for p in params]
        ctx.save_for_backward(low.detach().to(dtype).contiguous(), mid, flat)
        return out

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("the fused upscaler does not differentiate with respect to init_state_low "
                                      "(a constant tensor in every script, GS2D:616-619)")
        low, mid, flat = ctx.saved_tensors
        g = g.detach().to(ctx.geo.dtype).contiguous()
        gp = upscaler_bwd(ctx.geo, flat, low, mid, g.data_ptr())
        grads, o = [], 0
        for shp in ctx.shapes:
            n = 1
            for s in shp:
                n *= int(s)
            grads.append(gp[o:o + n].view(shp))
            o += n
        return (None, None, *grads)


class FusedUpscaler(nn.Module):
    """Base of the drop-in `upscaler` classes.  Subclasses register the reference's layers (so `state_dict` keys and
    initial values are the reference's) and name them in `_up_layers`, in packing order."""

    up_ndim = 2
    up_channels = 8
    up_act = "sigmoid"
    up_stride2 = 2
    _geo_cache = None

    def _up_modules(self) -> List[nn.Module]:
        raise NotImplementedError

    def up_parameters(self) -> List[torch.Tensor]:
        ps = []
        for m in self._up_modules():
            ps += [m.weight, m.bias]
        return ps

    def geometry(self, low_shape, dtype, device, **slab) -> UpscalerGeometry:
        key = (tuple(low_shape), dtype, str(device), tuple(sorted(slab.items())))
        if self._geo_cache is None:
            self._geo_cache = {}
        if key not in self._geo_cache:
            layers = len(self._up_modules()) - 1
            self._geo_cache[key] = UpscalerGeometry(self.up_ndim, self.up_channels, self.up_act, layers, self.up_stride2, low_shape,
                                                    dtype, device, **slab)
        return self._geo_cache[key]

    def forward(self, h: torch.Tensor) -> torch.Tensor:
        _require_cuda(h, "init_state_low")
        return _UpscalerFn.apply(self, h, *self.up_parameters())

    def scatter_flat_grad(self, flat_grad: torch.Tensor) -> None:
        """Adds a flat gradient (packing order) to the parameters' .grad (slab training: after the all-reduce)."""
        o = 0
        for p in self.up_parameters():
            n = p.numel()
            g = flat_grad[o:o + n].view(p.shape).to(p.dtype)
            p.grad = g.clone() if p.grad is None else p.grad + g
            o += n


class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred: torch.Tensor, target: torch.Tensor):
        _require_cuda(pred, "IC-loss prediction")
        p = pred.detach().contiguous()
        t = target.detach().to(p.dtype).contiguous()
        if p.shape != t.shape:
            raise ValueError(f"IC loss: prediction {tuple(p.shape)} vs target {tuple(t.shape)}")
        if p.dtype not in (torch.float32, torch.float64):
            raise TypeError(f"IC loss supports float32/float64, got {p.dtype}")
        L = _lib.lib()
        loss = torch.empty((), dtype=p.dtype, device=p.device)
        dt = _lib.F32 if p.dtype == torch.float32 else _lib.F64
        with torch.cuda.device(p.device):
            check(L.percnn_mse_fwd(dt, _device_index(p.device), p.data_ptr(), t.data_ptr(), p.numel(), loss.data_ptr(),
                                   _workspace(p.device, int(L.percnn_mse_workspace_bytes())).data_ptr(), _stream_ptr(p.device)))
        ctx.save_for_backward(p, t)
        return loss

    @staticmethod
    def backward(ctx, g_loss: torch.Tensor):
        p, t = ctx.saved_tensors
        g = torch.empty_like(p)
        gs = g_loss.detach().to(p.dtype).reshape(1).contiguous()
        dt = _lib.F32 if p.dtype == torch.float32 else _lib.F64
        with torch.cuda.device(p.device):
            check(_lib.lib().percnn_mse_bwd(dt, _device_index(p.device), p.data_ptr(), t.data_ptr(), p.numel(), gs.data_ptr(),
                                            g.data_ptr(), 0, _stream_ptr(p.device)))
        return g, None


def mse(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """`nn.MSELoss()(pred, target)` with a deterministic fp64 reduction; differentiable w.r.t. `pred`."""
    return _MSE.apply(pred, target)


def ic_target(low: torch.Tensor, mode: str, size: Tuple[int, ...]) -> torch.Tensor:
    """The constant the IC loss compares with: the interpolated low-resolution state (GS2D:334 'bicubic', GS3D:328
    'trilinear', BUR1:465-470 'bicubic_periodic' = periodic extension by one row/column, align_corners, crop).
    It does not depend on any parameter; plain ATen interpolation of a constant input."""
    with torch.no_grad():
        if mode == "bicubic_periodic":
            ext = torch.cat((low, low[:, :, :, 0:1]), dim=3)
            ext = torch.cat((ext, ext[:, :, 0:1, :]), dim=2)
            return F.interpolate(ext, tuple(n + 1 for n in size), mode="bicubic", align_corners=True)[:, :, :-1, :-1].contiguous()
        return F.interpolate(low, tuple(size), mode=mode).contiguous()


def ic_loss(model, mode: str, size: Optional[Tuple[int, ...]] = None) -> torch.Tensor:
    """`get_ic_loss(model)`: mse(model.UpconvBlock(model.init_state_low), interpolated init_state_low).  `size` defaults
    to the upscaler's own output size (the scripts hard-code it for their data: (100, 100), (48, 48, 48), (101, 101))."""
    low = model.init_state_low
    pred = model.UpconvBlock(low)
    if size is None:
        size = tuple(pred.shape[2:])
    return mse(pred, ic_target(low, mode, tuple(size)))
