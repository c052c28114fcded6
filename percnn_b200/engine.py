"""Host-side engine: plans, parameter packing and the autograd bridge to the C-ABI.

Everything here is plumbing (PyTorch owns device memory and streams); the arithmetic happens in
libpercnn_b200.so.  There is deliberately no CPU / eager fallback: a CPU tensor or a missing library
raises.
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import Desc, check


@dataclasses.dataclass(frozen=True)
class CellSpec:
    """The constants the reference hard-codes in each RCNNCell constructor."""
    cell: int                 # _lib.CELL_*
    ndim: int
    dtype: torch.dtype
    ksize: int = 1
    hidden: int = 0
    coef_mode: int = _lib.COEF_SIGMOID
    mu_up: float = 1.0
    dt: float = 1.0
    dx: float = 1.0
    flags: int = 0


@dataclasses.dataclass(frozen=True)
class DataLossSpec:
    """Which states and which points enter the fused data loss (percnn_data_loss_t).

    `mse_loss(output[0:-1:15, :, ::2, ::2, ::2], truth_sub)` (GS3D:403) over a rollout whose `output` holds every
    state is `DataLossSpec(sel=[s % 15 == 0 and s < nsteps for s in range(nsteps + 1)], stride=2)`.
    """
    sel: Tuple[bool, ...]      # nsteps + 1 entries; sel[s]: state h_s takes part
    stride: int = 1            # `::stride` on every spatial axis
    n_total: int = 0           # elements in the mean; 0 = this plan's own count (slab ranks pass the global one)

    @property
    def nsel(self) -> int:
        return sum(1 for e in self.sel if e)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"percnn_b200: {what} is on {t.device}; the fused cell runs on CUDA (sm_100a) only and has no CPU fallback")


class Plan:
    """percnn_plan_t wrapper bound to one (spec, spatial shape, device)."""

    def __init__(self, spec: CellSpec, spatial: Sequence[int], device: torch.device, slab_ghost: bool = False):
        L = _lib.lib()
        if len(spatial) != spec.ndim:
            raise ValueError(f"expected {spec.ndim} spatial dims, got {tuple(spatial)}")
        self.spec = spec
        self.spatial = tuple(int(s) for s in spatial)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("percnn_b200 plans need a CUDA device; there is no CPU fallback")
        self.slab_ghost = bool(slab_ghost)
        d = Desc()
        d.abi_version = _lib.ABI_VERSION
        d.ndim = spec.ndim
        ext = (1,) + self.spatial if spec.ndim == 2 else self.spatial
        for i in range(3):
            d.extent[i] = ext[i]
        d.dtype = _lib.F32 if spec.dtype == torch.float32 else _lib.F64
        d.cell, d.ksize, d.hidden, d.coef_mode, d.flags = spec.cell, spec.ksize, spec.hidden, spec.coef_mode, spec.flags
        d.mu_up, d.dt, d.dx = spec.mu_up, spec.dt, spec.dx
        d.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        d.slab_ghost = 1 if slab_ghost else 0
        self._h = ctypes.c_void_p()
        self._L = L
        with torch.cuda.device(self.device):
            check(L.percnn_plan_create(ctypes.byref(d), ctypes.byref(self._h)))
        self.nparams = int(L.percnn_param_count(self._h))
        self.state_elems = int(L.percnn_state_elems(self._h))
        self.uses_tma = bool(L.percnn_plan_uses_tma(self._h))
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        h = getattr(self, "_h", None)
        try:
            if h is not None and h.value:
                self._L.percnn_plan_destroy(h)
                h.value = None
        except Exception:       # interpreter shutdown: module globals may already be gone
            pass

    # -- shapes ---------------------------------------------------------------------------------
    @property
    def buffer_shape(self) -> Tuple[int, ...]:
        """Shape of one state buffer (2 fields; slowest axis carries the ghosts in slab mode)."""
        s = list(self.spatial)
        if self.slab_ghost:
            s[0] += 4
        return (2, *s)

    @property
    def tile2d_steps_per_pass(self) -> int:
        """> 0: 2-D rollouts run on shared-memory tiles, this many time steps per pass; 0: gather kernels."""
        return int(self._L.percnn_plan_uses_tile2d(self._h))

    @property
    def slab_persistent(self) -> bool:
        """Slab plans: rollouts of >= 2 steps run as one persistent cooperative kernel (small slabs)."""
        return bool(self._L.percnn_plan_slab_persistent(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.percnn_plan_launch_count(self._h))

    def workspace(self) -> torch.Tensor:
        if self._ws is None:
            nbytes = int(self._L.percnn_workspace_bytes(self._h, 0))
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def _check_state(self, t: torch.Tensor, what: str, slots: int = 1) -> None:
        _require_cuda(t, what)
        if t.dtype != self.spec.dtype:
            raise TypeError(f"{what}: dtype {t.dtype} != plan dtype {self.spec.dtype}")
        if not t.is_contiguous():
            raise ValueError(f"{what} must be contiguous")
        if t.numel() != slots * self.state_elems:
            raise ValueError(f"{what}: {t.numel()} elements, expected {slots} x {self.state_elems}")

    # -- calls ----------------------------------------------------------------------------------
    def params_load(self, flat: torch.Tensor) -> None:
        _require_cuda(flat, "params")
        if flat.dtype != self.spec.dtype or flat.numel() != self.nparams or not flat.is_contiguous():
            raise ValueError(f"params: need {self.nparams} contiguous {self.spec.dtype} scalars, got {flat.numel()} {flat.dtype}")
        check(self._L.percnn_params_load(self._h, flat.data_ptr(), _stream_ptr(self.device)))

    def step_fwd(self, h_in: torch.Tensor, h_out: torch.Tensor) -> None:
        self._check_state(h_in, "h_in")
        self._check_state(h_out, "h_out")
        check(self._L.percnn_step_fwd(self._h, h_in.data_ptr(), h_out.data_ptr(), _stream_ptr(self.device)))

    def step_rk4(self, h_in: torch.Tensor, h_out: torch.Tensor) -> None:
        self._check_state(h_in, "h_in")
        self._check_state(h_out, "h_out")
        check(self._L.percnn_step_rk4(self._h, h_in.data_ptr(), h_out.data_ptr(), self.workspace().data_ptr(),
                                      _stream_ptr(self.device)))

    def step_fwd_range(self, h_in: torch.Tensor, h_out: torch.Tensor, z_lo: int, z_hi: int) -> None:
        check(self._L.percnn_step_fwd_range(self._h, h_in.data_ptr(), h_out.data_ptr(), int(z_lo), int(z_hi),
                                            _stream_ptr(self.device)))

    def step_fwd_fused_halo(self, h_in: torch.Tensor, h_out: torch.Tensor, link) -> None:
        check(self._L.percnn_step_fwd_fused_halo(self._h, h_in.data_ptr(), h_out.data_ptr(), ctypes.byref(link),
                                                 _stream_ptr(self.device)))

    def step_bwd_fused_halo(self, h_in, g_out, g_in, link, g_add=None) -> None:
        check(self._L.percnn_step_bwd_fused_halo(self._h, h_in.data_ptr(), g_out.data_ptr(), _ptr(g_add), g_in.data_ptr(),
                                                 self.workspace().data_ptr(), ctypes.byref(link), _stream_ptr(self.device)))

    # -- whole slab rollouts (one C call each; see percnn_b200.halo) ---------------------------------
    def slab_rollout_fwd(self, ring, cur: int, nsteps: int, epoch: int) -> None:
        check(self._L.percnn_slab_rollout_fwd(self._h, ctypes.byref(ring), int(cur), int(nsteps), int(epoch) & 0xFFFFFFFF,
                                              _stream_ptr(self.device)))

    def slab_rollout_fwd_blocked(self, ring, wide, cur: int, nsteps: int, epoch: int) -> None:
        check(self._L.percnn_slab_rollout_fwd_blocked(self._h, ctypes.byref(ring), ctypes.byref(wide), int(cur), int(nsteps),
                                                      int(epoch) & 0xFFFFFFFF, _stream_ptr(self.device)))

    def slab_rollout_tape(self, tape, peer_lo_tape, peer_hi_tape, ring, nsteps: int, epoch: int) -> None:
        self._check_state(tape, "tape", nsteps + 1)
        check(self._L.percnn_slab_rollout_tape(self._h, tape.data_ptr(), peer_lo_tape.data_ptr(), peer_hi_tape.data_ptr(),
                                               ctypes.byref(ring), int(nsteps), int(epoch) & 0xFFFFFFFF,
                                               _stream_ptr(self.device)))

    def slab_rollout_bwd(self, tape, g_tape, spec, target, gscale, ring, nsteps: int, epoch: int) -> None:
        """Fused-halo adjoint over the whole tape; the parameter sums accumulate in the workspace header."""
        self._check_state(tape, "tape", nsteps + 1)
        dl_ref, keep = None, None
        if spec is not None:
            dl, keep = self._data_loss_struct(spec, nsteps, target, gscale)
            dl_ref = ctypes.byref(dl)
        check(self._L.percnn_slab_rollout_bwd(self._h, tape.data_ptr(), _ptr(g_tape), dl_ref, ctypes.byref(ring), int(nsteps),
                                              int(epoch) & 0xFFFFFFFF, self.workspace().data_ptr(), _stream_ptr(self.device)))
        del keep

    def reduction_sums(self, n: int = 24) -> torch.Tensor:
        """fp64 view of the running parameter-gradient sums at the head of the workspace (for the all-reduce)."""
        return self.workspace()[:8 * n].view(torch.float64)

    def step_bwd(self, h_in, g_out, g_in, g_add=None) -> None:
        for t, n in ((h_in, "h_in"), (g_out, "g_out"), (g_in, "g_in")):
            self._check_state(t, n)
        if g_add is not None:
            self._check_state(g_add, "g_add")
        check(self._L.percnn_step_bwd(self._h, h_in.data_ptr(), g_out.data_ptr(), _ptr(g_add), g_in.data_ptr(),
                                      self.workspace().data_ptr(), _stream_ptr(self.device)))

    def param_grads_begin(self) -> None:
        check(self._L.percnn_param_grads_begin(self._h, self.workspace().data_ptr(), _stream_ptr(self.device)))

    def param_grads_finish(self, flat: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(flat)
        check(self._L.percnn_param_grads_finish(self._h, flat.data_ptr(), out.data_ptr(), self.workspace().data_ptr(),
                                                _stream_ptr(self.device)))
        return out

    def rollout_fwd(self, h0: torch.Tensor, nsteps: int, *, tape: Optional[torch.Tensor] = None,
                    traj: Optional[torch.Tensor] = None, emit: Optional[Sequence[bool]] = None,
                    h_final: Optional[torch.Tensor] = None) -> None:
        self._check_state(h0, "h0")
        if tape is not None:
            self._check_state(tape, "tape", nsteps + 1)
        emit_arr = None
        if traj is not None:
            if emit is None or len(emit) != nsteps:
                raise ValueError("traj needs an emit mask with one entry per step")
            self._check_state(traj, "traj", sum(1 for e in emit if e))
            emit_arr = (ctypes.c_uint8 * nsteps)(*[1 if e else 0 for e in emit])
        if h_final is not None:
            self._check_state(h_final, "h_final")
        check(self._L.percnn_rollout_fwd(self._h, h0.data_ptr(), _ptr(traj), emit_arr, int(nsteps), _ptr(h_final),
                                         _ptr(tape), self.workspace().data_ptr(), _stream_ptr(self.device)))

    def rollout_bwd(self, flat: torch.Tensor, tape: torch.Tensor, g_tape: Optional[torch.Tensor],
                    gmask: Optional[Sequence[bool]], nsteps: int) -> Tuple[torch.Tensor, torch.Tensor]:
        self._check_state(tape, "tape", nsteps + 1)
        mask_arr = None
        if g_tape is not None:
            if gmask is None or len(gmask) != nsteps + 1:
                raise ValueError("g_tape needs a mask with nsteps+1 entries")
            self._check_state(g_tape, "g_tape", sum(1 for e in gmask if e))
            mask_arr = (ctypes.c_uint8 * (nsteps + 1))(*[1 if e else 0 for e in gmask])
        g_h0 = torch.empty(self.buffer_shape, dtype=self.spec.dtype, device=self.device)
        g_flat = torch.empty_like(flat)
        check(self._L.percnn_rollout_bwd(self._h, flat.data_ptr(), tape.data_ptr(), _ptr(g_tape), mask_arr, int(nsteps),
                                         g_h0.data_ptr(), g_flat.data_ptr(), self.workspace().data_ptr(),
                                         _stream_ptr(self.device)))
        return g_h0, g_flat

    # -- fused data loss ------------------------------------------------------------------------
    def lowres_shape(self, stride: int) -> Tuple[int, ...]:
        """Shape of one state's low-res target frame: (2, ceil(D/s), ceil(H/s), ceil(W/s)) (no D in 2-D)."""
        return (2, *[(n + stride - 1) // stride for n in self.spatial])

    def _data_loss_struct(self, spec: DataLossSpec, nsteps: int, target: torch.Tensor, gscale: Optional[torch.Tensor]):
        if len(spec.sel) != nsteps + 1:
            raise ValueError(f"data loss: selection mask needs nsteps + 1 = {nsteps + 1} entries, got {len(spec.sel)}")
        if spec.stride < 1:
            raise ValueError("data loss: stride must be >= 1")
        _require_cuda(target, "data-loss target")
        want = (spec.nsel, *self.lowres_shape(spec.stride))
        if target.dtype != self.spec.dtype or not target.is_contiguous() or tuple(target.shape) != want:
            raise ValueError(f"data-loss target: need a contiguous {self.spec.dtype} tensor of shape {want}, "
                             f"got {target.dtype} {tuple(target.shape)}")
        dl = _lib.DataLoss()
        sel_arr = (ctypes.c_uint8 * (nsteps + 1))(*[1 if e else 0 for e in spec.sel])
        dl.target, dl.sel, dl.stride, dl.reserved, dl.n_total = target.data_ptr(), sel_arr, int(spec.stride), 0, int(spec.n_total)
        if gscale is not None:
            _require_cuda(gscale, "data-loss gradient scale")
            if gscale.numel() != 1 or gscale.dtype != self.spec.dtype:
                raise ValueError("data-loss gradient scale must be one scalar of the plan dtype")
            dl.gscale = gscale.data_ptr()
        return dl, sel_arr          # sel_arr must outlive the call

    def data_loss_fwd(self, tape: torch.Tensor, nsteps: int, spec: DataLossSpec, target: torch.Tensor) -> torch.Tensor:
        """0-dim tensor: sum over selected states / sampled points of (h - target)^2, divided by n_total."""
        self._check_state(tape, "tape", nsteps + 1)
        dl, _keep = self._data_loss_struct(spec, nsteps, target, None)
        out = torch.empty((), dtype=self.spec.dtype, device=self.device)
        check(self._L.percnn_data_loss_fwd(self._h, tape.data_ptr(), int(nsteps), ctypes.byref(dl), out.data_ptr(),
                                           self.workspace().data_ptr(), _stream_ptr(self.device)))
        return out

    def rollout_bwd_loss(self, flat: torch.Tensor, tape: torch.Tensor, nsteps: int, spec: DataLossSpec,
                         target: torch.Tensor, gscale: Optional[torch.Tensor] = None,
                         g_tape: Optional[torch.Tensor] = None, gmask: Optional[Sequence[bool]] = None):
        """rollout_bwd with the loss gradient injected by the adjoint kernels (no dense gradient tape)."""
        self._check_state(tape, "tape", nsteps + 1)
        dl, _keep = self._data_loss_struct(spec, nsteps, target, gscale)
        mask_arr = None
        if g_tape is not None:
            if gmask is None or len(gmask) != nsteps + 1:
                raise ValueError("g_tape needs a mask with nsteps+1 entries")
            self._check_state(g_tape, "g_tape", sum(1 for e in gmask if e))
            mask_arr = (ctypes.c_uint8 * (nsteps + 1))(*[1 if e else 0 for e in gmask])
        g_h0 = torch.empty(self.buffer_shape, dtype=self.spec.dtype, device=self.device)
        g_flat = torch.empty_like(flat)
        check(self._L.percnn_rollout_bwd_loss(self._h, flat.data_ptr(), tape.data_ptr(), _ptr(g_tape), mask_arr,
                                              ctypes.byref(dl), int(nsteps), g_h0.data_ptr(), g_flat.data_ptr(),
                                              self.workspace().data_ptr(), _stream_ptr(self.device)))
        return g_h0, g_flat

    def step_bwd_loss(self, h_in, g_out, g_in, *, target_frame=None, stride=1, n_total=0, gscale=None, g_add=None,
                      link=None) -> None:
        """One adjoint step with the loss gradient of state h_in injected (link: fused-halo slab step)."""
        check(self._L.percnn_step_bwd_loss(self._h, h_in.data_ptr(), g_out.data_ptr(), _ptr(g_add), _ptr(target_frame),
                                           int(stride), int(n_total), _ptr(gscale), g_in.data_ptr(),
                                           self.workspace().data_ptr(), None if link is None else ctypes.byref(link),
                                           _stream_ptr(self.device)))

    def rollout_fwd_host(self, flat_host: torch.Tensor, h0_host: torch.Tensor, nsteps: int,
                         emit: Optional[Sequence[bool]] = None, want_final: bool = True):
        """End-to-end call with HOST buffers (pinned or pageable): H2D, rollout, D2H, sync."""
        if flat_host.is_cuda or h0_host.is_cuda:
            raise ValueError("rollout_fwd_host takes host tensors")
        nemit = 0 if emit is None else sum(1 for e in emit if e)
        # page-locking a GiB-sized buffer costs ~100 ms, so the pinned result buffers are kept across calls
        # (the caller must consume/copy a result before the next call overwrites it)
        cache = self.__dict__.setdefault("_host_out", {})
        traj = fin = None
        if nemit:
            traj = cache.get(("traj", nemit))
            if traj is None:
                traj = cache[("traj", nemit)] = torch.empty((nemit, *self.buffer_shape), dtype=self.spec.dtype).pin_memory()
        if want_final:
            fin = cache.get("fin")
            if fin is None:
                fin = cache["fin"] = torch.empty(self.buffer_shape, dtype=self.spec.dtype).pin_memory()
        emit_arr = None if emit is None else (ctypes.c_uint8 * nsteps)(*[1 if e else 0 for e in emit])
        check(self._L.percnn_rollout_fwd_host(self._h, flat_host.data_ptr(), h0_host.data_ptr(), _ptr(traj), emit_arr,
                                              int(nsteps), _ptr(fin)))
        return traj, fin


_PLAN_CACHE: Dict[tuple, Plan] = {}
_PLAN_ORDER: List[tuple] = []
_MAX_PLANS = 5  # the library has 6 constant-memory parameter slots


def get_plan(spec: CellSpec, spatial: Sequence[int], device: torch.device, slab_ghost: bool = False) -> Plan:
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (spec, tuple(int(s) for s in spatial), str(device), bool(slab_ghost))
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        while len(_PLAN_ORDER) >= _MAX_PLANS:
            old = _PLAN_ORDER.pop(0)
            _PLAN_CACHE.pop(old, None)
        plan = Plan(spec, spatial, device, slab_ghost)
        _PLAN_CACHE[key] = plan
        _PLAN_ORDER.append(key)
    return plan


def clear_plans() -> None:
    _PLAN_CACHE.clear()
    _PLAN_ORDER.clear()


def pack_params(tensors: Sequence[torch.Tensor], dtype: torch.dtype) -> torch.Tensor:
    """Flat packing = the cell's state_dict tensors concatenated in state_dict order."""
    return torch.cat([t.detach().reshape(-1).to(dtype) for t in tensors]).contiguous()


class _Rollout(torch.autograd.Function):
    """states[0] = h0, states[s+1] = cell(states[s]);  backward = the hand-derived adjoint kernels.

    Only the states are stored (8 B per cell per step in fp32) -- the reference's autograd keeps
    146-705 B per cell per step.
    """

    @staticmethod
    def forward(ctx, plan: Plan, nsteps: int, h0: torch.Tensor, *params: torch.Tensor):
        _require_cuda(h0, "state")
        flat = pack_params(params, plan.spec.dtype)
        plan.params_load(flat)
        states = torch.empty((nsteps + 1, *plan.buffer_shape), dtype=plan.spec.dtype, device=h0.device)
        plan.rollout_fwd(h0.detach().contiguous().view(plan.buffer_shape), nsteps, tape=states)
        ctx.plan, ctx.nsteps = plan, nsteps
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.dtypes = [p.dtype for p in params]
        ctx.save_for_backward(states, flat)
        return states

    @staticmethod
    def backward(ctx, g_states: torch.Tensor):
        states, flat = ctx.saved_tensors
        plan, nsteps = ctx.plan, ctx.nsteps
        g_states = g_states.contiguous()
        plan.params_load(flat)
        g_h0, g_flat = plan.rollout_bwd(flat, states, g_states, [True] * (nsteps + 1), nsteps)
        grads: List[Optional[torch.Tensor]] = []
        off = 0
        for i, (shape, dt) in enumerate(zip(ctx.shapes, ctx.dtypes)):
            n = 1
            for s in shape:
                n *= s
            if ctx.needs_input_grad[3 + i]:
                grads.append(g_flat[off:off + n].view(shape).to(dt))
            else:
                grads.append(None)
            off += n
        return (None, None, g_h0 if ctx.needs_input_grad[2] else None, *grads)


def _split_param_grads(ctx, g_flat: torch.Tensor, first: int) -> List[Optional[torch.Tensor]]:
    grads: List[Optional[torch.Tensor]] = []
    off = 0
    for i, (shape, dt) in enumerate(zip(ctx.shapes, ctx.dtypes)):
        n = 1
        for s in shape:
            n *= s
        grads.append(g_flat[off:off + n].view(shape).to(dt) if ctx.needs_input_grad[first + i] else None)
        off += n
    return grads


class _RolloutLoss(torch.autograd.Function):
    """(states, loss) = rollout + fused strided-subsample MSE.

    The loss value is one small reduction over the sampled points of the tape; in backward its gradient is
    injected by the adjoint kernels themselves, so no dense [nsteps+1, 2, ...] gradient is ever materialised
    unless the caller ALSO differentiates through `states` (then both sources are combined).
    """

    @staticmethod
    def forward(ctx, plan: Plan, nsteps: int, spec: DataLossSpec, target: torch.Tensor, h0: torch.Tensor,
                *params: torch.Tensor):
        _require_cuda(h0, "state")
        flat = pack_params(params, plan.spec.dtype)
        plan.params_load(flat)
        states = torch.empty((nsteps + 1, *plan.buffer_shape), dtype=plan.spec.dtype, device=h0.device)
        plan.rollout_fwd(h0.detach().contiguous().view(plan.buffer_shape), nsteps, tape=states)
        target = target.detach().contiguous()
        loss = plan.data_loss_fwd(states, nsteps, spec, target)
        ctx.plan, ctx.nsteps, ctx.spec = plan, nsteps, spec
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.dtypes = [p.dtype for p in params]
        ctx.save_for_backward(states, flat, target)
        ctx.set_materialize_grads(False)
        return states, loss

    @staticmethod
    def backward(ctx, g_states: Optional[torch.Tensor], g_loss: Optional[torch.Tensor]):
        states, flat, target = ctx.saved_tensors
        plan, nsteps = ctx.plan, ctx.nsteps
        none = (None,) * (5 + len(ctx.shapes))
        if g_states is None and g_loss is None:
            return none
        plan.params_load(flat)
        dense = None if g_states is None else g_states.contiguous()
        mask = None if dense is None else [True] * (nsteps + 1)
        if g_loss is None:
            g_h0, g_flat = plan.rollout_bwd(flat, states, dense, mask, nsteps)
        else:
            gscale = g_loss.detach().to(plan.spec.dtype).reshape(1).contiguous()
            g_h0, g_flat = plan.rollout_bwd_loss(flat, states, nsteps, ctx.spec, target, gscale, dense, mask)
        return (None, None, None, None, g_h0 if ctx.needs_input_grad[4] else None, *_split_param_grads(ctx, g_flat, 5))


def rollout_states_with_data_loss(plan: Plan, nsteps: int, h0: torch.Tensor, params: Sequence[torch.Tensor],
                                  spec: DataLossSpec, target: torch.Tensor):
    """(states [nsteps+1, 2, ...], loss 0-dim); both differentiable w.r.t. h0 and the parameters."""
    return _RolloutLoss.apply(plan, nsteps, spec, target, h0, *params)


def rollout_states(plan: Plan, nsteps: int, h0: torch.Tensor, params: Sequence[torch.Tensor]) -> torch.Tensor:
    """All states of an nsteps rollout as one [nsteps+1, 2, ...] tensor (differentiable)."""
    return _Rollout.apply(plan, nsteps, h0, *params)


@torch.no_grad()
def rollout_emit(plan: Plan, nsteps: int, h0: torch.Tensor, params: Sequence[torch.Tensor],
                 emit: Sequence[bool], want_final: bool = False):
    """Inference rollout that stores only the emitted states (ping-pong scratch for the rest)."""
    _require_cuda(h0, "state")
    flat = pack_params(params, plan.spec.dtype)
    plan.params_load(flat)
    nemit = sum(1 for e in emit if e)
    traj = torch.empty((nemit, *plan.buffer_shape), dtype=plan.spec.dtype, device=h0.device)
    fin = torch.empty(plan.buffer_shape, dtype=plan.spec.dtype, device=h0.device) if want_final else None
    plan.rollout_fwd(h0.contiguous().view(plan.buffer_shape), nsteps, traj=traj if nemit else None,
                     emit=emit if nemit else None, h_final=fin)
    return traj, fin
