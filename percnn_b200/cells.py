"""Host-side mirror of the reference's `RCNNCell` / `RCNN` nn.Module surface.

Same constructor signatures, attribute names, parameter names/shapes and state_dict keys as the
reference scripts, so their `train()`, checkpoint and post-processing code runs unchanged; `forward`
dispatches to the fused CUDA kernels through the C-ABI instead of issuing ~35 ATen ops per step.

One generic implementation is configured per variant in `percnn_b200.variants.*`:

    from percnn_b200.variants.gs3d import RCNNCell, RCNN, upscaler      # DataDrivenModeling/3d_gs_rd
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib, engine
from .engine import CellSpec

# 4th-order central-difference tables (GS2D:20-24, GS3D:22-39, BUR3:20-36), built rather than typed in
_LAP_1D = (-1.0 / 12.0, 4.0 / 3.0, -5.0 / 2.0, 4.0 / 3.0, -1.0 / 12.0)
_DER_1D = (1.0 / 12.0, -8.0 / 12.0, 0.0, 8.0 / 12.0, -1.0 / 12.0)


def laplace_table(ndim: int) -> np.ndarray:
    t = np.zeros((1, 1) + (5,) * ndim)
    for ax in range(ndim):
        for o, w in zip(range(-2, 3), _LAP_1D):
            idx = [2] * ndim
            idx[ax] += o
            t[(0, 0) + tuple(idx)] += w
    return t


def derivative_table(axis: int) -> np.ndarray:
    """dx_2d_op (axis 0 = rows, BUR3:20-24) / dy_2d_op (axis 1 = columns, BUR3:26-30)."""
    t = np.zeros((1, 1, 5, 5))
    if axis == 0:
        t[0, 0, :, 2] = _DER_1D
    else:
        t[0, 0, 2, :] = _DER_1D
    return t


def _conv_nd(ndim):
    return nn.Conv2d if ndim == 2 else nn.Conv3d


class _FusedCell(nn.Module):
    """Common machinery: parameter packing, plan lookup, the step and the rollout."""

    ndim: int = 2
    dtype: torch.dtype = torch.float32

    def _spec(self) -> CellSpec:  # pragma: no cover - overridden
        raise NotImplementedError

    def _packed_tensors(self) -> List[torch.Tensor]:
        """state_dict order == the C-ABI packing order."""
        return list(self.state_dict(keep_vars=True).values())

    def _plan(self, h: torch.Tensor) -> engine.Plan:
        if h.dim() != self.ndim + 2 or h.shape[0] != 1 or h.shape[1] != 2:
            raise ValueError(f"expected a state of shape [1, 2, {'D, ' if self.ndim == 3 else ''}H, W], got {tuple(h.shape)}")
        if h.dtype != self.dtype:
            raise TypeError(f"state dtype {h.dtype} does not match the cell's {self.dtype}")
        return engine.get_plan(self._spec(), tuple(h.shape[2:]), h.device)

    def forward(self, h: torch.Tensor):
        """One explicit-Euler step; returns the new state twice, like the reference (GS2D:119-121)."""
        plan = self._plan(h)
        states = engine.rollout_states(plan, 1, h[0], self._packed_tensors())
        ch = states[1:2]
        return ch, ch

    def rollout(self, h0: torch.Tensor, nsteps: int) -> torch.Tensor:
        """[nsteps+1, 2, ...] tensor of states (slot 0 = h0); differentiable w.r.t. h0 and parameters."""
        plan = self._plan(h0)
        return engine.rollout_states(plan, int(nsteps), h0[0], self._packed_tensors())

    def rollout_data_loss(self, h0: torch.Tensor, nsteps: int, target: torch.Tensor, sel: Sequence[bool], stride: int):
        """(states, loss): the rollout plus the fused strided-subsample MSE over the states with sel[s] set,
        `mse_loss(states[sel][:, :, ::stride, ::stride(, ::stride)], target)` -- the data loss of the training
        scripts (GS3D:403, GS2D:397-401, BUR1:610-614).  Its gradient is injected inside the adjoint kernels, so
        backward needs no dense gradient of `states`."""
        plan = self._plan(h0)
        spec = engine.DataLossSpec(sel=tuple(bool(e) for e in sel), stride=int(stride))
        return engine.rollout_states_with_data_loss(plan, int(nsteps), h0[0], self._packed_tensors(), spec, target)

    def rollout_emit(self, h0: torch.Tensor, nsteps: int, emit: Sequence[bool], want_final: bool = False):
        plan = self._plan(h0)
        return engine.rollout_emit(plan, int(nsteps), h0[0], self._packed_tensors(), emit, want_final)

    def init_hidden_tensor(self, prev_state):
        return prev_state.cuda()


class PiCell(_FusedCell):
    """Pi-block cell: q+ = q + dt (alpha_q Lap q + Wh4_q(Wh1_q h * Wh2_q h * Wh3_q h)).

    Parameter names, shapes and registration order follow GS2D:61-86 / GS3D:76-101 / BUR1:99-124 /
    FWD:42-71.
    """

    def _build(self, *, ndim, dtype, ksize, hidden, dx, dt, coef_mode, mu_up, coef_names, coef_init,
               init_scale, init_kind, flags=0):
        self.ndim, self.dtype = ndim, dtype
        self._ksize, self._hidden, self._coef_mode, self._flags = ksize, hidden, coef_mode, flags
        self.dx, self.dt = dx, dt
        self._mu = mu_up
        for name, val in zip(coef_names, coef_init):
            setattr(self, name, nn.Parameter(torch.tensor(val, dtype=dtype), requires_grad=True))
        Conv = _conv_nd(ndim)
        self.W_laplace = Conv(1, 1, 5, 1, padding=0, bias=False)
        lap = torch.tensor(laplace_table(ndim), dtype=dtype)
        # fp32 scripts scale as 1/dx**2 * table (GS2D:66), the fp64 one as table / dx**2 (FWD:48)
        self.W_laplace.weight.data = (lap / dx ** 2) if dtype == torch.float64 else (1 / dx ** 2 * lap)
        self.W_laplace.weight.requires_grad = False
        self.filter_list = []
        for q in "uv":
            for i in (1, 2, 3):
                conv = Conv(2, hidden, ksize, 1, padding=0, bias=True).to(dtype)
                setattr(self, f"Wh{i}_{q}", conv)
                self.filter_list.append(conv)
            conv = Conv(hidden, 1, 1, 1, padding=0, bias=True).to(dtype)
            setattr(self, f"Wh4_{q}", conv)
            self.filter_list.append(conv)
        self._init_kind = init_kind
        self.init_filter(self.filter_list, c=init_scale)

    def init_filter(self, filter_list, c):
        """Xavier * c (GS2D:92-103) or U(+-c sqrt(1/prod(shape[:-1]))) (FWD:86-95, BUR1:126-135)."""
        for f in filter_list:
            if self._init_kind == "xavier":
                nn.init.xavier_uniform_(f.weight)
                f.weight.data = c * f.weight.data
            else:
                bound = c * np.sqrt(1 / np.prod(f.weight.shape[:-1]))
                f.weight.data.uniform_(-bound, bound)
            if f.bias is not None:
                f.bias.data.fill_(0.0)

    def _spec(self) -> CellSpec:
        return CellSpec(cell=_lib.CELL_PI, ndim=self.ndim, dtype=self.dtype, ksize=self._ksize, hidden=self._hidden,
                        coef_mode=self._coef_mode, mu_up=float(self._mu), dt=float(self.dt), dx=float(self.dx),
                        flags=self._flags)


class Conv2dDerivative(nn.Module):
    """Fixed circular finite-difference filter divided by its resolution (BUR3:54-80).

    Kept as a real module so `laplace_op.filter.weight` etc. appear in the state_dict; the fused
    kernels read the taps from it, `forward` is only used off the hot path.
    """

    def __init__(self, DerFilter, resol, kernel_size=5, name=""):
        super().__init__()
        self.resol = resol
        self.name = name
        self.input_channels = self.output_channels = 1
        self.kernel_size = kernel_size
        self.input_padding = self.padding = kernel_size // 2
        self.filter = nn.Conv2d(1, 1, kernel_size, 1, padding=self.padding, padding_mode="circular", bias=False)
        self.filter.weight.data = torch.tensor(DerFilter, dtype=torch.float64)
        self.filter.weight.requires_grad = False

    def forward(self, input):
        return self.filter(input) / self.resol


class PhysicsCell(_FusedCell):
    """Stage-3 cells: Euler step of a closed-form RHS with trainable scalar coefficients."""

    dtype = torch.float64

    def f_rhs(self, u, v):
        """Reference RHS through stock convs (off the hot path; the fused kernel implements the same)."""
        raise NotImplementedError

    def _rk4_stock(self, h):
        """The reference's formulation through stock ops (BUR3:159-206): only used to differentiate forward_rk4."""
        u0, v0 = h[:, 0:1, ...], h[:, 1:2, ...]
        k1u, k1v = self.f_rhs(u0, v0)
        k2u, k2v = self.f_rhs(u0 + k1u * self.dt / 2.0, v0 + k1v * self.dt / 2.0)
        k3u, k3v = self.f_rhs(u0 + k2u * self.dt / 2.0, v0 + k2v * self.dt / 2.0)
        k4u, k4v = self.f_rhs(u0 + k3u * self.dt, v0 + k3v * self.dt)
        return torch.cat((u0 + self.dt * (k1u + 2 * k2u + 2 * k3u + k4u) / 6.0,
                          v0 + self.dt * (k1v + 2 * k2v + 2 * k3v + k4v) / 6.0), dim=1)

    def forward_rk4(self, h):
        """Classical RK4 on f_rhs (BUR3:159-206, LO3:153-200) -- defined by the reference, never called by its scripts.
        Forward: four launches of the fused right-hand side (percnn_step_rk4).  Backward (no reference script trains
        through it): autograd of the stock-op formulation, recomputed from the saved input."""
        plan = self._plan(h)
        ch = _RK4Step.apply(self, plan, h, *self._packed_tensors())
        return ch, ch


class _RK4Step(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cell, plan, h, *params):
        flat = engine.pack_params(params, plan.spec.dtype)
        plan.params_load(flat)
        out = torch.empty(plan.buffer_shape, dtype=plan.spec.dtype, device=h.device)
        plan.step_rk4(h.detach()[0].contiguous(), out)
        ctx.cell = cell
        ctx.params = params
        ctx.save_for_backward(h.detach())
        return out[None]

    @staticmethod
    def backward(ctx, g):
        (h,) = ctx.saved_tensors
        cell = ctx.cell
        wanted = [i for i, p in enumerate(ctx.params) if ctx.needs_input_grad[3 + i]]
        with torch.enable_grad():
            hh = h.clone().requires_grad_(True)
            out = cell._rk4_stock(hh)
            grads = torch.autograd.grad(out, [hh] + [ctx.params[i] for i in wanted], g, allow_unused=True)
        pg = [None] * len(ctx.params)
        for i, gr in zip(wanted, grads[1:]):
            pg[i] = gr
        return (None, None, grads[0] if ctx.needs_input_grad[2] else None, *pg)


def data_loss_selection(step: int, effective_step, time_stride: int, first_frames: Optional[int] = None):
    """Which states `torch.cat(outputs)[0:-1:time_stride]` picks (GS3D:394-403): `outputs` is [h_0] followed by
    h_{s+1} for every s < step in `effective_step` (GS3D:191-212), the slice drops the last entry and keeps every
    time_stride-th one; `first_frames` mirrors `pred[:idx]` (GS2D:398-401).  Returns (frame_state, sel) with
    outputs[i] = state frame_state[i] and sel[s] = state s enters the loss (step + 1 entries).  Pure host logic."""
    if int(time_stride) < 1:
        raise ValueError("time_stride must be >= 1")
    eff = set(int(s) for s in effective_step)
    frame_state = [0] + [s + 1 for s in range(int(step)) if s in eff]
    picked = [frame_state[i] for i in range(0, len(frame_state) - 1, int(time_stride))]
    if first_frames is not None:
        picked = picked[:int(first_frames)]
    if not picked:
        raise ValueError("data loss selects no frame")
    sel = [False] * (int(step) + 1)
    for st in picked:
        sel[st] = True
    return frame_state, sel


class FusedRCNN(nn.Module):
    """`RCNN.forward()` (GS2D:162-190): unroll `step` cell steps, collect the effective ones.

    The whole unroll is ONE library call (`percnn_rollout_fwd`) writing every state into a single
    [step+1, 2, ...] tensor; the returned list entries are views of it, so `torch.cat(outputs)` and
    the caller's slicing work as before and autograd flows through the hand-written adjoint.
    """

    cell_attr = "crnn_cell"

    def _setup(self, cell: _FusedCell, step, effective_step):
        self.step = step
        self.effective_step = effective_step
        self._all_layers = []
        setattr(self, self.cell_attr, cell)
        self._all_layers.append(cell)

    def _initial_state(self) -> torch.Tensor:
        return self.UpconvBlock(self.init_state_low)

    def forward(self):
        self.init_state = self._initial_state()
        cell: _FusedCell = getattr(self, self.cell_attr)
        outputs = [self.init_state]
        second_last_state = []
        if self.step <= 0:
            return outputs, second_last_state
        eff = set(int(s) for s in self.effective_step)
        if torch.is_grad_enabled() or len(eff) * 2 >= self.step:
            states = cell.rollout(self.init_state, self.step)
            for s in range(self.step):
                if s in eff:
                    outputs.append(states[s + 1:s + 2])
            if self.step >= 2:
                second_last_state = states[self.step - 1:self.step].clone()
        else:
            # inference with sparse emission: keep only what the caller will see
            emit = [(s in eff) or (s == self.step - 2) for s in range(self.step)]
            traj, _ = cell.rollout_emit(self.init_state, self.step, emit)
            slot = 0
            for s in range(self.step):
                if emit[s]:
                    if s in eff:
                        outputs.append(traj[slot:slot + 1])
                    if s == self.step - 2:
                        second_last_state = traj[slot:slot + 1].clone()
                    slot += 1
        return outputs, second_last_state

    def forward_data_loss(self, truth_sub: torch.Tensor, time_stride: int, space_stride: int):
        """`forward()` plus the scripts' data loss in one pass: returns (outputs, second_last_state, loss_data) with

            loss_data == mse_loss(torch.cat(outputs)[0:-1:time_stride, :, ::space_stride, ...], truth_sub)

        (GS3D:394-403, GS2D:394-401, BUR1:606-614).  The selection of frames follows the reference exactly
        (list index -> state index through `effective_step`); the loss and its gradient are computed by the fused
        kernels (`percnn_data_loss_fwd`, injection inside the adjoint), not by slicing a dense trajectory."""
        self.init_state = self._initial_state()
        cell: _FusedCell = getattr(self, self.cell_attr)
        frame_state, sel = data_loss_selection(self.step, self.effective_step, time_stride)
        states, loss = cell.rollout_data_loss(self.init_state, self.step, truth_sub, sel, int(space_stride))
        outputs = [self.init_state] + [states[st:st + 1] for st in frame_state[1:]]
        second_last_state = states[self.step - 1:self.step].clone() if self.step >= 2 else []
        return outputs, second_last_state, loss
