"""CPU oracle for the PeRCNN recurrent-cell hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference's `RCNNCell.forward`
/ `RCNN.forward` for every variant on the hot path.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl
reference` legs may import it; the product package `percnn_b200` never does.

Parity status: PINNED against outputs of the reference itself.  The reference
has no tests and no golden vectors of its own, so
`tests/golden/make_golden.py` imports the reference's own `RCNNCell`/`RCNN`
classes from /root/reference (in the build container), runs them with the
shipped checkpoints' weights and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this file against those vectors.

Two independent restatements live here:

* `*_torch`  -- the same ATen op sequence the reference issues (cat-padding,
  conv2d/conv3d, mul, sigmoid, add), so it tracks the reference to the last bit
  on the same machine and is what the CPU baseline times ("port" of the
  reference's CPU PyTorch path; the only third-party code on that path is
  PyTorch itself: nn.Conv2d / nn.Conv3d / torch.cat / torch.sigmoid).
* `*_np`     -- a direct numpy stencil written from the maths (np.roll for the
  periodic shifts, explicit tap sums, fp64 by default) that shares no code with
  the first; used as the high-precision yardstick.

Reference aliases used in the citations (paths under /root/reference):
  FWD   ForwardSimulationOfPDEs/2d_lambda_omega/percnn_LO_eqn.py
  GS2D  DataDrivenModeling/2d_gs_rd/train_2drd.py
  GS3D  DataDrivenModeling/3d_gs_rd/train_3drd.py
  BUR1  DataDrivenDiscoveryOfPDEs/2D_Burgers_eqn/Stage-1/rcnn_Burgers_[...].py
  LO1   DataDrivenDiscoveryOfPDEs/2D_Lambda_Omega_eqn/stage-1/rcnn_LO_[...].py
  BUR3  DataDrivenDiscoveryOfPDEs/2D_Burgers_eqn/Stage-3/fine_tuning_[5%noise,41x51x51].py
  LO3   DataDrivenDiscoveryOfPDEs/2D_Lambda_Omega_eqn/stage-3/fine_tuning_LO_[0%noise,41x51x51].py
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Fixed finite-difference stencils (radius 2, 4th-order central differences).
# --------------------------------------------------------------------------------------

#: 1-D second-derivative taps at offsets -2..+2; the 2-D / 3-D Laplacians are the sum of
#: this stencil along every axis (centre = ndim * -5/2).  GS2D:20-24, GS3D:22-39.
LAP_1D = (-1.0 / 12.0, 4.0 / 3.0, -5.0 / 2.0, 4.0 / 3.0, -1.0 / 12.0)
#: 1-D first-derivative taps at offsets -2..+2.  BUR3:20-30.
DER_1D = (1.0 / 12.0, -8.0 / 12.0, 0.0, 8.0 / 12.0, -1.0 / 12.0)


def laplace_stencil(ndim: int) -> np.ndarray:
    """Dense [1,1,5,5(,5)] Laplacian table exactly as the reference builds it.

    2-D: `lap_2d_op` GS2D:20-24 / `lapl_op` FWD:18-22; 3-D: `laplace_3d` GS3D:22-39.
    """
    st = np.zeros((1, 1) + (5,) * ndim)
    c = (2,) * ndim
    st[(0, 0) + c] = -5.0 if ndim == 2 else -15.0 / 2.0
    for ax in range(ndim):
        for off, w in ((-2, -1.0 / 12.0), (-1, 4.0 / 3.0), (1, 4.0 / 3.0), (2, -1.0 / 12.0)):
            idx = list(c)
            idx[ax] += off
            st[(0, 0) + tuple(idx)] = w
    return st


def dx_stencil_2d() -> np.ndarray:
    """`dx_2d_op` BUR3:20-24 -- taps along tensor dim 2 (rows)."""
    st = np.zeros((1, 1, 5, 5))
    st[0, 0, :, 2] = DER_1D
    return st


def dy_stencil_2d() -> np.ndarray:
    """`dy_2d_op` BUR3:26-30 -- taps along tensor dim 3 (columns)."""
    st = np.zeros((1, 1, 5, 5))
    st[0, 0, 2, :] = DER_1D
    return st


# --------------------------------------------------------------------------------------
# Variant table: the constants the reference hard-codes inside each RCNNCell constructor.
# --------------------------------------------------------------------------------------


@dataclasses.dataclass(frozen=True)
class Variant:
    name: str
    ndim: int
    dtype: torch.dtype
    kind: str  # "pi" | "burgers" | "lo"
    k: int  # Pi-block conv kernel size (1 or 5); 0 for physics cells
    hc: int  # Pi-block hidden channels
    dx: float
    dt: float
    coef_mode: str  # "raw" (FWD: DA*Lap) | "sigmoid" (mu_up*sigmoid(CA)*Lap)
    mu_up: float
    coef_names: Tuple[str, ...]
    cite: str


VARIANTS: Dict[str, Variant] = {
    v.name: v
    for v in (
        Variant("fwd", 2, torch.float64, "pi", 1, 4, 0.2, 0.0125, "raw", 1.0, ("DA", "DB"), "FWD:24-112"),
        Variant("gs2d", 2, torch.float32, "pi", 1, 8, 0.01, 0.5, "sigmoid", 3.99e-5, ("CA", "CB"), "GS2D:43-121"),
        Variant("gs3d", 3, torch.float32, "pi", 1, 2, 100 / 48, 0.5, "sigmoid", 0.274, ("CA", "CB"), "GS3D:58-139"),
        Variant("bur1", 2, torch.float32, "pi", 5, 16, 1 / 100, 0.00025, "sigmoid", 0.01, ("CA", "CB"), "BUR1:54-178"),
        Variant("lo1", 2, torch.float32, "pi", 5, 16, 0.2, 0.0125, "sigmoid", 0.2, ("CA", "CB"), "LO1:53-171"),
        Variant("bur3", 2, torch.float64, "burgers", 0, 0, 1 / 100, 0.00025, "raw", 1.0,
                ("nu_u", "nu_v", "C1_u", "C2_u", "C1_v", "C2_v"), "BUR3:83-221"),
        Variant("lo3", 2, torch.float64, "lo", 0, 0, 0.2, 0.0125, "raw", 1.0,
                ("nu_u", "nu_v", "C1_u", "C2_u", "C3_u", "C4_u", "C5_u",
                 "C1_v", "C2_v", "C3_v", "C4_v", "C5_v", "C6_v"), "LO3:83-215"),
    )
}

PI_CONV_NAMES = ("Wh1_u", "Wh2_u", "Wh3_u", "Wh4_u", "Wh1_v", "Wh2_v", "Wh3_v", "Wh4_v")

#: Literal initial coefficients of the Stage-3 scripts (BUR3:123-130, LO3:123-136; C6_v is the
#: extra term of the 10%-noise twin, 0 reproduces the 0%-noise script).
BUR3_LITERALS = dict(nu_u=0.0050078, nu_v=0.0050228, C1_u=-0.982252, C2_u=-0.992132,
                     C1_v=-0.983758, C2_v=-0.971269)
LO3_LITERALS = dict(nu_u=0.09465, nu_v=0.09455, C1_u=1.0081, C2_u=-1.0167, C3_u=0.9973,
                    C4_u=-1.0176, C5_u=0.9981, C1_v=0.9873, C2_v=-0.9987, C3_v=-0.9945,
                    C4_v=-0.9985, C5_v=-0.9928, C6_v=0.0)


def make_pi_params(variant: str, seed: int = 0, hc: Optional[int] = None,
                   dtype: Optional[torch.dtype] = None, scale: float = 0.5) -> Dict[str, torch.Tensor]:
    """Seeded random parameter set with the reference's names and shapes.

    The distribution is NOT the reference's init (Xavier*c, GS2D:92-103); tests want
    weights large enough that every term of the update matters.
    """
    v = VARIANTS[variant]
    hc = v.hc if hc is None else hc
    dtype = v.dtype if dtype is None else dtype
    g = torch.Generator().manual_seed(seed)
    ksz = (v.k,) * v.ndim
    one = (1,) * v.ndim
    p: Dict[str, torch.Tensor] = {}
    for n in v.coef_names:
        p[n] = (torch.rand((), generator=g, dtype=torch.float64) - 0.3).to(dtype)
    p["W_laplace.weight"] = (torch.tensor(laplace_stencil(v.ndim), dtype=dtype) / v.dx ** 2
                             if dtype == torch.float64 else
                             1 / v.dx ** 2 * torch.tensor(laplace_stencil(v.ndim), dtype=dtype))
    for q in "uv":
        for i in (1, 2, 3):
            w = (torch.rand((hc, 2) + ksz, generator=g, dtype=torch.float64) - 0.5) * 2 * scale
            if v.k > 1:
                w = w / v.k
            p[f"Wh{i}_{q}.weight"] = w.to(dtype)
            p[f"Wh{i}_{q}.bias"] = ((torch.rand((hc,), generator=g, dtype=torch.float64) - 0.5) * scale).to(dtype)
        p[f"Wh4_{q}.weight"] = ((torch.rand((1, hc) + one, generator=g, dtype=torch.float64) - 0.5) * 2 * scale).to(dtype)
        p[f"Wh4_{q}.bias"] = ((torch.rand((1,), generator=g, dtype=torch.float64) - 0.5) * scale).to(dtype)
    return p


def make_phys_params(variant: str, dtype: Optional[torch.dtype] = None,
                     jitter_seed: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Stage-3 coefficient set = the scripts' literals, optionally jittered."""
    v = VARIANTS[variant]
    dtype = v.dtype if dtype is None else dtype
    lit = BUR3_LITERALS if v.kind == "burgers" else LO3_LITERALS
    g = None if jitter_seed is None else torch.Generator().manual_seed(jitter_seed)
    p = {}
    for n in v.coef_names:
        val = torch.tensor(lit[n], dtype=torch.float64)
        if g is not None:
            val = val + 0.05 * (torch.rand((), generator=g, dtype=torch.float64) - 0.5)
        p[n] = val.to(dtype)
    return p


# --------------------------------------------------------------------------------------
# Restatement 1: the reference's ATen op sequence (torch, CPU).
# --------------------------------------------------------------------------------------


def periodic_pad(h: torch.Tensor, r: int = 2) -> torch.Tensor:
    """Manual wrap-around padding by concatenation, last axis first.  GS2D:108-109, GS3D:125-127."""
    nd = h.dim() - 2
    for ax in range(h.dim() - 1, h.dim() - 1 - nd, -1):
        n = h.shape[ax]
        h = torch.cat((h.narrow(ax, n - r, r), h, h.narrow(ax, 0, r)), dim=ax)
    return h


def _conv(x, w, b=None):
    return F.conv2d(x, w, b) if w.dim() == 4 else F.conv3d(x, w, b)


def pi_cell_step_torch(h: torch.Tensor, p: Dict[str, torch.Tensor], v: Variant) -> torch.Tensor:
    """One explicit-Euler step of a Pi-block cell.  SURVEY Appendix A.

    GS2D:105-121 (k=1, 2-D), FWD:98-112 (raw DA/DB), GS3D:123-139 (3-D), BUR1:161-176 (k=5:
    the Pi convs read the +-2 padded state with padding=0); LO1:165-166 pads circularly inside
    the convs, which is the same arithmetic.
    """
    h_pad = periodic_pad(h, 2)
    out = []
    for qi, q in enumerate("uv"):
        q_pad = h_pad[:, qi:qi + 1]
        q_prev = h[:, qi:qi + 1]
        c = p[v.coef_names[qi]]
        alpha = c if v.coef_mode == "raw" else v.mu_up * torch.sigmoid(c)
        pi_in = h_pad if v.k == 5 else h
        prod = (_conv(pi_in, p[f"Wh1_{q}.weight"], p[f"Wh1_{q}.bias"])
                * _conv(pi_in, p[f"Wh2_{q}.weight"], p[f"Wh2_{q}.bias"])
                * _conv(pi_in, p[f"Wh3_{q}.weight"], p[f"Wh3_{q}.bias"]))
        res = alpha * _conv(q_pad, p["W_laplace.weight"]) + _conv(prod, p[f"Wh4_{q}.weight"], p[f"Wh4_{q}.bias"])
        out.append(q_prev + res * v.dt)
    return torch.cat(out, dim=1)


def _circ_conv2d(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d(..., padding=2, padding_mode='circular', bias=False).  BUR3:70-72."""
    return F.conv2d(F.pad(x, (2, 2, 2, 2), mode="circular"), w)


def phys_rhs_torch(u: torch.Tensor, vv: torch.Tensor, p: Dict[str, torch.Tensor], v: Variant):
    """`f_rhs` of a Stage-3 physics cell: Burgers BUR3:154-157, lambda-omega LO3:148-151.  The derivative filters
    hold the un-scaled taps and the result is divided by `resol` afterwards (BUR3:78-80)."""
    dt_ = u.dtype
    lap_w = torch.tensor(laplace_stencil(2), dtype=dt_)
    lap = lambda x: _circ_conv2d(x, lap_w) / (v.dx ** 2)
    if v.kind == "burgers":
        dxw = torch.tensor(dx_stencil_2d(), dtype=dt_)
        dyw = torch.tensor(dy_stencil_2d(), dtype=dt_)
        ddx = lambda x: _circ_conv2d(x, dxw) / v.dx
        ddy = lambda x: _circ_conv2d(x, dyw) / v.dx
        f_u = p["nu_u"] * lap(u) + p["C1_u"] * u * ddx(u) + p["C2_u"] * vv * ddy(u)
        f_v = p["nu_v"] * lap(vv) + p["C1_v"] * u * ddx(vv) + p["C2_v"] * vv * ddy(vv)
    else:
        f_u = (p["nu_u"] * lap(u) + p["C1_u"] * u + p["C2_u"] * u ** 3 + p["C3_u"] * u ** 2 * vv
               + p["C4_u"] * u * vv ** 2 + p["C5_u"] * vv ** 3)
        f_v = (p["nu_v"] * lap(vv) + p["C1_v"] * vv + p["C2_v"] * u ** 3 + p["C3_v"] * u ** 2 * vv
               + p["C4_v"] * u * vv ** 2 + p["C5_v"] * vv ** 3)
        if "C6_v" in p:
            f_v = f_v + p["C6_v"] * u
    return f_u, f_v


def phys_cell_step_torch(h: torch.Tensor, p: Dict[str, torch.Tensor], v: Variant) -> torch.Tensor:
    """One Euler step of a Stage-3 physics cell: `f_rhs` + `forward` (BUR3:209-221, LO3:203-215)."""
    u, vv = h[:, 0:1], h[:, 1:2]
    f_u, f_v = phys_rhs_torch(u, vv, p, v)
    return torch.cat((u + v.dt * f_u, vv + v.dt * f_v), dim=1)


def rk4_step_torch(h: torch.Tensor, p: Dict[str, torch.Tensor], variant: str) -> torch.Tensor:
    """`RCNNCell.forward_rk4` (BUR3:159-206, LO3:153-200): classical RK4 on `f_rhs`, same operation order."""
    v = VARIANTS[variant]
    u0, v0 = h[:, 0:1], h[:, 1:2]
    k1u, k1v = phys_rhs_torch(u0, v0, p, v)
    k2u, k2v = phys_rhs_torch(u0 + k1u * v.dt / 2.0, v0 + k1v * v.dt / 2.0, p, v)
    k3u, k3v = phys_rhs_torch(u0 + k2u * v.dt / 2.0, v0 + k2v * v.dt / 2.0, p, v)
    k4u, k4v = phys_rhs_torch(u0 + k3u * v.dt, v0 + k3v * v.dt, p, v)
    return torch.cat((u0 + v.dt * (k1u + 2 * k2u + 2 * k3u + k4u) / 6.0,
                      v0 + v.dt * (k1v + 2 * k2v + 2 * k3v + k4v) / 6.0), dim=1)


def cell_step_torch(h: torch.Tensor, p: Dict[str, torch.Tensor], variant: str) -> torch.Tensor:
    v = VARIANTS[variant]
    return pi_cell_step_torch(h, p, v) if v.kind == "pi" else phys_cell_step_torch(h, p, v)


def rollout_torch(h0: torch.Tensor, p: Dict[str, torch.Tensor], variant: str, step: int,
                  effective_step: Sequence[int]) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """`RCNN.forward` GS2D:162-190: outputs = [h0] + [state after step s for s in effective_step];
    second_last_state = clone of the state after step index `step-2`."""
    eff = set(int(s) for s in effective_step)
    h = h0
    outputs = [h0]
    second_last = []
    for s in range(step):
        h = cell_step_torch(h, p, variant)
        if s == step - 2:
            second_last = h.clone()
        if s in eff:
            outputs.append(h)
    return outputs, second_last


# --------------------------------------------------------------------------------------
# Restatement 2: direct numpy stencil, independent of the conv formulation.
# --------------------------------------------------------------------------------------


def _np(t, dtype=np.float64):
    return t.detach().cpu().numpy().astype(dtype) if isinstance(t, torch.Tensor) else np.asarray(t, dtype=dtype)


def _shift(a: np.ndarray, off: int, axis: int) -> np.ndarray:
    """a shifted so that result[i] = a[i + off] with periodic wrap."""
    return np.roll(a, -off, axis=axis)


def _apply_taps_np(q: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Cross-correlate a dense 5^n tap table with a periodic field (skips zero taps)."""
    nd = q.ndim
    out = np.zeros_like(q)
    for idx in np.argwhere(w != 0):
        sh = q
        for ax in range(nd):
            sh = _shift(sh, int(idx[ax]) - 2, ax)
        out += w[tuple(idx)] * sh
    return out


def cell_step_np(h, p, variant: str, dtype=np.float64) -> np.ndarray:
    """Same step as `cell_step_torch`, written from the maths of SURVEY 2.4:

        P_i^q = W_i^q * h + b_i^q ; R^q = sum_c W_4^q[c] (P_1 P_2 P_3)[c] + b_4^q
        L^q = Lap * q ; q+ = q + dt (alpha_q L^q + R^q)
    """
    v = VARIANTS[variant]
    h = _np(h, dtype)[0]  # [2, ...]
    u, vv = h[0], h[1]
    nd = v.ndim
    if v.kind == "pi":
        lapw = _np(p["W_laplace.weight"], dtype)[0, 0]
        out = []
        for qi, q in enumerate("uv"):
            c = float(_np(p[v.coef_names[qi]], dtype))
            alpha = c if v.coef_mode == "raw" else v.mu_up / (1.0 + np.exp(-c))
            L = _apply_taps_np(h[qi], lapw)
            prod = 1.0
            for i in (1, 2, 3):
                W = _np(p[f"Wh{i}_{q}.weight"], dtype)
                b = _np(p[f"Wh{i}_{q}.bias"], dtype)
                hc, k = W.shape[0], W.shape[2]
                r = k // 2
                P = np.zeros((hc,) + h[qi].shape, dtype=dtype)
                for c_ in range(hc):
                    acc = np.full(h[qi].shape, b[c_], dtype=dtype)
                    for f in range(2):
                        for idx in np.ndindex(*W.shape[2:]):
                            sh = h[f]
                            for ax in range(nd):
                                sh = _shift(sh, idx[ax] - r, ax)
                            acc = acc + W[(c_, f) + idx] * sh
                    P[c_] = acc
                prod = prod * P
            W4 = _np(p[f"Wh4_{q}.weight"], dtype).reshape(-1)
            b4 = float(_np(p[f"Wh4_{q}.bias"], dtype).reshape(-1)[0])
            R = np.tensordot(W4, prod, axes=(0, 0)) + b4
            out.append(h[qi] + v.dt * (alpha * L + R))
        return np.stack(out)[None]
    lapw = laplace_stencil(2)[0, 0] / v.dx ** 2
    lap = lambda a: _apply_taps_np(a, lapw)
    g = lambda n: float(_np(p[n], dtype))
    if v.kind == "burgers":
        ddx = lambda a: _apply_taps_np(a, dx_stencil_2d()[0, 0] / v.dx)
        ddy = lambda a: _apply_taps_np(a, dy_stencil_2d()[0, 0] / v.dx)
        f_u = g("nu_u") * lap(u) + g("C1_u") * u * ddx(u) + g("C2_u") * vv * ddy(u)
        f_v = g("nu_v") * lap(vv) + g("C1_v") * u * ddx(vv) + g("C2_v") * vv * ddy(vv)
    else:
        f_u = (g("nu_u") * lap(u) + g("C1_u") * u + g("C2_u") * u ** 3 + g("C3_u") * u ** 2 * vv
               + g("C4_u") * u * vv ** 2 + g("C5_u") * vv ** 3)
        f_v = (g("nu_v") * lap(vv) + g("C1_v") * vv + g("C2_v") * u ** 3 + g("C3_v") * u ** 2 * vv
               + g("C4_v") * u * vv ** 2 + g("C5_v") * vv ** 3)
        if "C6_v" in p:
            f_v = f_v + g("C6_v") * u
    return np.stack((u + v.dt * f_u, vv + v.dt * f_v))[None]


# --------------------------------------------------------------------------------------
# Hand-derived adjoint, numpy fp64 -- the yardstick for the backward kernels.
# --------------------------------------------------------------------------------------


def cell_step_vjp_np(h, g_out, p, variant: str):
    """Vector-Jacobian product of one step: returns (g_in, {param: grad}).

    Formulas of SURVEY 8a ("Hand-derived adjoint for a1-a3" and the a4/a5 blocks), restated with
    np.roll; Lap^T = Lap, Dx^T = -Dx, Dy^T = -Dy under the periodic boundary.
    """
    v = VARIANTS[variant]
    dtype = np.float64
    h = _np(h, dtype)[0]
    G = _np(g_out, dtype)[0]
    nd = v.ndim
    u, vv = h[0], h[1]
    grads: Dict[str, np.ndarray] = {}
    g_in = G.copy()
    if v.kind == "pi":
        lapw = _np(p["W_laplace.weight"], dtype)[0, 0]
        for qi, q in enumerate("uv"):
            Gq = G[qi]
            c = float(_np(p[v.coef_names[qi]], dtype))
            sig = 1.0 / (1.0 + np.exp(-c))
            alpha = c if v.coef_mode == "raw" else v.mu_up * sig
            L = _apply_taps_np(h[qi], lapw)
            gL = v.dt * np.sum(Gq * L)
            grads[v.coef_names[qi]] = np.array(gL if v.coef_mode == "raw" else gL * v.mu_up * sig * (1 - sig))
            g_in[qi] += v.dt * alpha * _apply_taps_np(Gq, lapw)
            Ws = [_np(p[f"Wh{i}_{q}.weight"], dtype) for i in (1, 2, 3)]
            bs = [_np(p[f"Wh{i}_{q}.bias"], dtype) for i in (1, 2, 3)]
            hc, k = Ws[0].shape[0], Ws[0].shape[2]
            r = k // 2
            shifts = {}
            for f in range(2):
                for idx in np.ndindex(*Ws[0].shape[2:]):
                    sh = h[f]
                    for ax in range(nd):
                        sh = _shift(sh, idx[ax] - r, ax)
                    shifts[(f,) + idx] = sh
            P = []
            for i in range(3):
                Pi = np.zeros((hc,) + u.shape)
                for c_ in range(hc):
                    acc = np.full(u.shape, bs[i][c_])
                    for key, sh in shifts.items():
                        acc = acc + Ws[i][(c_,) + key] * sh
                    Pi[c_] = acc
                P.append(Pi)
            W4 = _np(p[f"Wh4_{q}.weight"], dtype).reshape(-1)
            prod = P[0] * P[1] * P[2]
            grads[f"Wh4_{q}.bias"] = np.array([v.dt * Gq.sum()])
            grads[f"Wh4_{q}.weight"] = (v.dt * (Gq[None] * prod).reshape(hc, -1).sum(1)).reshape(
                _np(p[f"Wh4_{q}.weight"]).shape)
            for i in range(3):
                others = [P[j] for j in range(3) if j != i]
                Gbar = v.dt * Gq[None] * W4.reshape((hc,) + (1,) * nd) * others[0] * others[1]
                grads[f"Wh{i + 1}_{q}.bias"] = Gbar.reshape(hc, -1).sum(1)
                gW = np.zeros_like(Ws[i])
                for key, sh in shifts.items():
                    gW[(slice(None),) + key] = (Gbar * sh[None]).reshape(hc, -1).sum(1)
                    # dL/dh_f(x) += sum_c W[c,f,a] * Gbar[c](x - a + r)
                    f = key[0]
                    contrib = np.tensordot(Ws[i][(slice(None),) + key], Gbar, axes=(0, 0))
                    for ax in range(nd):
                        contrib = _shift(contrib, -(key[1 + ax] - r), ax)
                    g_in[f] += contrib
                grads[f"Wh{i + 1}_{q}.weight"] = gW
        return g_in[None], grads
    lapw = laplace_stencil(2)[0, 0] / v.dx ** 2
    lap = lambda a: _apply_taps_np(a, lapw)
    gp = lambda n: float(_np(p[n], dtype))
    Gu, Gv = G[0], G[1]
    dt = v.dt
    if v.kind == "burgers":
        Dx = lambda a: _apply_taps_np(a, dx_stencil_2d()[0, 0] / v.dx)
        Dy = lambda a: _apply_taps_np(a, dy_stencil_2d()[0, 0] / v.dx)
        g_in[0] += dt * (gp("nu_u") * lap(Gu) + gp("C1_u") * Dx(u) * Gu - Dx(gp("C1_u") * u * Gu)
                         - Dy(gp("C2_u") * vv * Gu) + gp("C1_v") * Dx(vv) * Gv)
        g_in[1] += dt * (gp("nu_v") * lap(Gv) - Dx(gp("C1_v") * u * Gv) + gp("C2_v") * Dy(vv) * Gv
                         - Dy(gp("C2_v") * vv * Gv) + gp("C2_u") * Dy(u) * Gu)
        grads["nu_u"] = np.array(dt * np.sum(Gu * lap(u)))
        grads["nu_v"] = np.array(dt * np.sum(Gv * lap(vv)))
        grads["C1_u"] = np.array(dt * np.sum(Gu * u * Dx(u)))
        grads["C2_u"] = np.array(dt * np.sum(Gu * vv * Dy(u)))
        grads["C1_v"] = np.array(dt * np.sum(Gv * u * Dx(vv)))
        grads["C2_v"] = np.array(dt * np.sum(Gv * vv * Dy(vv)))
    else:
        c6 = gp("C6_v") if "C6_v" in p else 0.0
        dfu_du = gp("C1_u") + 3 * gp("C2_u") * u ** 2 + 2 * gp("C3_u") * u * vv + gp("C4_u") * vv ** 2
        dfu_dv = gp("C3_u") * u ** 2 + 2 * gp("C4_u") * u * vv + 3 * gp("C5_u") * vv ** 2
        dfv_du = 3 * gp("C2_v") * u ** 2 + 2 * gp("C3_v") * u * vv + gp("C4_v") * vv ** 2 + c6
        dfv_dv = gp("C1_v") + gp("C3_v") * u ** 2 + 2 * gp("C4_v") * u * vv + 3 * gp("C5_v") * vv ** 2
        g_in[0] += dt * (gp("nu_u") * lap(Gu) + dfu_du * Gu + dfv_du * Gv)
        g_in[1] += dt * (gp("nu_v") * lap(Gv) + dfu_dv * Gu + dfv_dv * Gv)
        grads["nu_u"] = np.array(dt * np.sum(Gu * lap(u)))
        grads["nu_v"] = np.array(dt * np.sum(Gv * lap(vv)))
        mon = {"2": u ** 3, "3": u ** 2 * vv, "4": u * vv ** 2, "5": vv ** 3}
        grads["C1_u"] = np.array(dt * np.sum(Gu * u))
        grads["C1_v"] = np.array(dt * np.sum(Gv * vv))
        for n, m in mon.items():
            grads[f"C{n}_u"] = np.array(dt * np.sum(Gu * m))
            grads[f"C{n}_v"] = np.array(dt * np.sum(Gv * m))
        if "C6_v" in p:
            grads["C6_v"] = np.array(dt * np.sum(Gv * u))
    return g_in[None], grads


# --------------------------------------------------------------------------------------
# Data loss of the training scripts and its gradient, restated in numpy
# --------------------------------------------------------------------------------------


def data_loss_frames(step: int, effective_step: Sequence[int], time_stride: int,
                     first_frames: Optional[int] = None) -> List[int]:
    """State indices that `torch.cat(outputs)[0:-1:time_stride]` picks (GS3D:394-403, GS2D:394-397).

    `outputs` = [h_0] + [h_{s+1} for s in range(step) if s in effective_step] (GS3D:191-212); the slice
    drops the last list entry (the dummy step) and keeps every time_stride-th one.  `first_frames`
    mirrors `pred[:idx]` of GS2D:398-401 / BUR1:611-614 (training part of the selected frames)."""
    eff = set(int(e) for e in effective_step)
    frame_state = [0] + [s + 1 for s in range(step) if s in eff]
    picked = [frame_state[i] for i in range(0, len(frame_state) - 1, time_stride)]
    return picked if first_frames is None else picked[:first_frames]


def data_loss_np(states, target, frames: Sequence[int], stride: int) -> float:
    """nn.MSELoss()(output[frames][:, :, ::s, ::s(, ::s)], target)  (GS3D:401-403), fp64.

    states: [nsteps+1, 2, ...]; target: [len(frames), 2, ceil(./s) ...]."""
    st = np.asarray(states, dtype=np.float64)
    nd = st.ndim - 2
    sub = st[list(frames)][(slice(None), slice(None)) + (slice(None, None, stride),) * nd]
    d = sub - np.asarray(target, dtype=np.float64)
    return float(np.mean(d * d))


def data_loss_grad_np(states, target, frames: Sequence[int], stride: int, gscale: float = 1.0) -> np.ndarray:
    """d(gscale * data_loss)/d(states): 2/N (h - target) on the sampling lattice of the picked states, 0 elsewhere."""
    st = np.asarray(states, dtype=np.float64)
    nd = st.ndim - 2
    sl = (slice(None), slice(None)) + (slice(None, None, stride),) * nd
    sub = st[list(frames)][sl]
    d = sub - np.asarray(target, dtype=np.float64)
    g = np.zeros_like(st)
    for i, f in enumerate(frames):
        g[f][sl[1:]] = (2.0 * gscale / d.size) * d[i]
    return g


# --------------------------------------------------------------------------------------
# Physics-residual loss of the training scripts, restated
# --------------------------------------------------------------------------------------

#: PDE constants hard-coded in each script's get_phy_Loss and loss_generator defaults
PHYS_CONSTS = {
    "fwd": dict(kind="lo", D=(0.1, 0.1), dt=0.0125, dx=0.2),                                  # FWD:268, 337-340
    "gs2d": dict(kind="gs", D=(2e-5, 2e-5 / 4), f=1 / 25, k=3 / 50, dt=1.0 / 2, dx=1.0 / 100),  # GS2D:244, 321-328
    "gs3d": dict(kind="gs", D=(0.2, 0.1), f=0.025, k=0.055, dt=0.5, dx=100 / 48),             # GS3D:267, 319-326
}


def _phys_reaction(kind, c, u, v):
    if kind == "lo":
        a = u * u + v * v
        return (1 - a) * u + a * v, -a * u + (1 - a) * v
    return -u * v * v + c["f"] * (1 - u), u * v * v - (c["f"] + c["k"]) * v


def phys_loss_torch(output: torch.Tensor, variant: str) -> torch.Tensor:
    """`loss_gen(output, loss_generator())` (FWD:288-357 and siblings) as the same sequence of steps on the
    un-padded trajectory [T, 2, ...]: periodic pad (2 before, 3 after), valid Laplacian of frames 0..T-3 on
    extent+1 points, forward time difference, PDE residual, mse(f_u, 0) + mse(f_v, 0).  Differentiable."""
    c = PHYS_CONSTS[variant]
    nd = output.dim() - 2
    pad = output
    for ax in range(2 + nd - 1, 1, -1):     # FWD:349-350: last axis first
        n = pad.shape[ax]
        pad = torch.cat((pad.narrow(ax, n - 2, 2), pad, pad.narrow(ax, 0, 3)), dim=ax)
    w = torch.tensor(laplace_stencil(nd), dtype=output.dtype) / c["dx"] ** 2
    conv = torch.nn.functional.conv2d if nd == 2 else torch.nn.functional.conv3d
    inner = (slice(None), slice(None)) + (slice(2, -2),) * nd
    lap_u = conv(pad[0:-2, 0:1], w)
    lap_v = conv(pad[0:-2, 1:2], w)
    q = pad[inner]
    q_t = (q[1:-1] - q[:-2]) / c["dt"]          # Conv1d([-1, 1, 0]) / dt over time (FWD:283-286, 324-327)
    u, v = q[0:-2, 0:1], q[0:-2, 1:2]
    ru, rv = _phys_reaction(c["kind"], c, u, v)
    f_u = c["D"][0] * lap_u + ru - q_t[:, 0:1]
    f_v = c["D"][1] * lap_v + rv - q_t[:, 1:2]
    return (f_u ** 2).mean() + (f_v ** 2).mean()


def phys_loss_np(output, variant: str):
    """Independent numpy/fp64 version written from the maths on the PERIODIC grid: residual by np.roll, the
    (2, 3) padding's double counting as the weight w(x) = prod_axes (1 + [x_axis == 0]).  Returns
    (loss, dloss/doutput)."""
    c = PHYS_CONSTS[variant]
    o = np.asarray(output, dtype=np.float64)
    T, nd = o.shape[0], o.ndim - 2
    taps = np.array([-1 / 12, 4 / 3, -5 / 2, 4 / 3, -1 / 12]) / c["dx"] ** 2

    def lap(a):      # a: [t, ...spatial]
        r = np.zeros_like(a)
        for ax in range(1, 1 + nd):
            for off, wgt in zip(range(-2, 3), taps):
                r += wgt * np.roll(a, -off, axis=ax)
        return r

    u, v = o[:-2, 0], o[:-2, 1]
    ru, rv = _phys_reaction(c["kind"], c, u, v)
    fu = c["D"][0] * lap(u) + ru - (o[1:-1, 0] - u) / c["dt"]
    fv = c["D"][1] * lap(v) + rv - (o[1:-1, 1] - v) / c["dt"]
    w = np.ones(o.shape[2:])
    n = T - 2
    for ax, ext in enumerate(o.shape[2:]):
        n *= ext + 1
        idx = [slice(None)] * nd
        idx[ax] = 0
        w[tuple(idx)] *= 2
    loss = float((w * (fu ** 2 + fv ** 2)).sum() / n)
    # gradient: R = 2 w f / N; d/dq_t = D Lap(R_q[t]) + J^T R[t] + R_q[t]/dt - R_q[t-1]/dt
    Ru, Rv = 2 * w * fu / n, 2 * w * fv / n
    if c["kind"] == "lo":
        a = u * u + v * v
        ruu, ruv = (1 - a) - 2 * u * u + 2 * u * v, -2 * u * v + a + 2 * v * v
        rvu, rvv = -a - 2 * u * u - 2 * u * v, -2 * u * v + (1 - a) - 2 * v * v
    else:
        ruu, ruv = -v * v - c["f"], -2 * u * v
        rvu, rvv = v * v, 2 * u * v - (c["f"] + c["k"])
    g = np.zeros_like(o)
    g[:-2, 0] += c["D"][0] * lap(Ru) + ruu * Ru + rvu * Rv + Ru / c["dt"]
    g[:-2, 1] += c["D"][1] * lap(Rv) + ruv * Ru + rvv * Rv + Rv / c["dt"]
    g[1:-1, 0] -= Ru / c["dt"]
    g[1:-1, 1] -= Rv / c["dt"]
    return loss, g


# --------------------------------------------------------------------------------------
# Stage-2 library of candidate terms
# --------------------------------------------------------------------------------------

STAGE2_LIST_A = ["ones", "u", "v", "u**2", "u*v", "v**2", "u**3", "u**2*v", "u*v**2", "v**3"]   # PDE_FIND_u.py:187
STAGE2_LIST_B = ["ones", "u_x", "u_y", "v_x", "v_y", "lap_u", "lap_v"]                           # PDE_FIND_u.py:188


def stage2_library_torch(output: torch.Tensor, kind: str, dt: float, dx: float) -> Dict[str, torch.Tensor]:
    """`Loss_generator.get_phy_residual` (BUR2d:129-199; lambda-omega: `get_library`) on the periodic (2, 3) padding
    `get_residual_mse` builds (BUR2d:207-208), as the same op sequence: three fixed 5x5 valid convs per field divided
    by their resolution, forward time difference, residual.  `output`: un-padded [T, 2, H, W]."""
    dtype = output.dtype
    pad = torch.cat((output[:, :, :, -2:], output, output[:, :, :, 0:3]), dim=3)
    pad = torch.cat((pad[:, :, -2:, :], pad, pad[:, :, 0:3, :]), dim=2)
    w_dx = torch.tensor(dx_stencil_2d(), dtype=torch.float32).to(dtype).reshape(1, 1, 5, 5)
    w_dy = torch.tensor(dy_stencil_2d(), dtype=torch.float32).to(dtype).reshape(1, 1, 5, 5)
    w_lap = torch.tensor(laplace_stencil(2), dtype=torch.float32).to(dtype).reshape(1, 1, 5, 5)
    u_in, v_in = pad[0:-2, 0:1], pad[0:-2, 1:2]
    lap_u, lap_v = F.conv2d(u_in, w_lap) / dx ** 2, F.conv2d(v_in, w_lap) / dx ** 2
    u_x, u_y = F.conv2d(u_in, w_dx) / dx, F.conv2d(u_in, w_dy) / dx
    v_x, v_y = F.conv2d(v_in, w_dx) / dx, F.conv2d(v_in, w_dy) / dx
    q = pad[:, :, 2:-2, 2:-2]
    q_t = (q[1:-1] - q[:-2]) / dt
    u, v, u_t, v_t = q[0:-2, 0:1], q[0:-2, 1:2], q_t[:, 0:1], q_t[:, 1:2]
    if kind == "burgers":
        nu = 1 / 200
        f_u = u_t - nu * lap_u + u * u_x + v * u_y
        f_v = v_t - nu * lap_v + u * v_x + v * v_y
    else:
        f_u = u_t - (0.1 * lap_u + (1 - u ** 2 - v ** 2) * u + 1.0 * (u ** 2 + v ** 2) * v)
        f_v = v_t - (0.1 * lap_v + (1 - u ** 2 - v ** 2) * v - 1.0 * (u ** 2 + v ** 2) * u)
    return {"f_u": f_u, "f_v": f_v, "ones": torch.ones_like(u), "u": u, "v": v, "u_t": u_t, "v_t": v_t, "u_x": u_x,
            "u_y": u_y, "v_x": v_x, "v_y": v_y, "lap_u": lap_u, "lap_v": lap_v}


def stage2_theta_np(library: Dict[str, torch.Tensor], idx) -> Tuple[np.ndarray, np.ndarray]:
    """PDE_FIND_u.py:228-259: terms -> flattened fp64 columns, the 70 products A*B at the sampled rows, rhs (u_t, v_t)."""
    col = {k: v.detach().double().numpy().reshape(-1)[np.asarray(idx)] for k, v in library.items()}
    u, v = col["u"], col["v"]
    A = {"ones": np.ones_like(u), "u": u, "v": v, "u**2": u ** 2, "u*v": u * v, "v**2": v ** 2, "u**3": u ** 3,
         "u**2*v": u ** 2 * v, "u*v**2": u * v ** 2, "v**3": v ** 3}
    lhs = np.stack([A[a] * col[b] for a in STAGE2_LIST_A for b in STAGE2_LIST_B], axis=1)
    return lhs, np.stack((col["u_t"], col["v_t"]), axis=1)


# --------------------------------------------------------------------------------------
# Initial-state generator ("upscaler") and IC loss
# --------------------------------------------------------------------------------------

#: per script family: (ndim, channels, activation, layers, stride of the 2nd transposed conv, state_dict key prefixes)
UPSCALERS = {
    "gs2d": dict(ndim=2, C=8, act="sigmoid", layers=2, stride2=2, keys=("convnet.0", "convnet.2", "convnet.3")),   # GS2D:26-41
    "gs3d": dict(ndim=3, C=8, act="sigmoid", layers=2, stride2=1, keys=("convnet.0", "convnet.2", "convnet.3")),   # GS3D:41-56
    "stage": dict(ndim=2, C=16, act="tanh", layers=1, stride2=1, keys=("up0", "out")),                            # BUR1:38-52
}


def upscaler_torch(low: torch.Tensor, sd: Dict[str, torch.Tensor], kind: str) -> torch.Tensor:
    """`upscaler.forward` as the ATen op sequence `nn.Sequential` issues (GS2D:40-41): conv_transpose (k5, s2, p2,
    op1), sigmoid | tanh, [conv_transpose (k5, s, p2, op s-1)], 1x1 conv.  Differentiable (CPU autograd gives the
    parameter gradients the fused adjoint is checked against)."""
    u = UPSCALERS[kind]
    ct = F.conv_transpose2d if u["ndim"] == 2 else F.conv_transpose3d
    cv = F.conv2d if u["ndim"] == 2 else F.conv3d
    k = u["keys"]
    x = ct(low, sd[k[0] + ".weight"], sd[k[0] + ".bias"], stride=2, padding=2, output_padding=1)
    x = torch.sigmoid(x) if u["act"] == "sigmoid" else torch.tanh(x)
    if u["layers"] == 2:
        s2 = u["stride2"]
        x = ct(x, sd[k[1] + ".weight"], sd[k[1] + ".bias"], stride=s2, padding=2, output_padding=s2 - 1)
    return cv(x, sd[k[-1] + ".weight"], sd[k[-1] + ".bias"])


def _conv_transpose_np(x: np.ndarray, w: np.ndarray, b: np.ndarray, stride: int) -> np.ndarray:
    """Transposed conv (kernel 5, padding 2, output_padding stride-1) from its definition, scatter form:
    out[co, s*i - 2 + k] += x[ci, i] * w[ci, co, k] per axis; x [Cin, *sp], w [Cin, Cout, 5, ...]."""
    nd = x.ndim - 1
    sp = x.shape[1:]
    full = tuple(stride * (n - 1) + 5 for n in sp)               # un-cropped scatter target
    out = np.zeros((w.shape[1],) + full, dtype=np.float64)
    for k in np.ndindex(*(5,) * nd):
        contrib = np.tensordot(w[(slice(None), slice(None)) + k].T.astype(np.float64), x.astype(np.float64), axes=(1, 0))
        sl = tuple(slice(kk, kk + stride * (n - 1) + 1, stride) for kk, n in zip(k, sp))
        out[(slice(None),) + sl] += contrib
    crop = tuple(slice(2, 2 + stride * n) for n in sp)           # padding 2 in front; extent stride*n (op = stride-1)
    res = np.zeros((w.shape[1],) + tuple(stride * n for n in sp), dtype=np.float64)
    src = out[(slice(None),) + crop]
    res[(slice(None),) + tuple(slice(0, m) for m in src.shape[1:])] = src
    return res + b.astype(np.float64).reshape((-1,) + (1,) * nd)


def upscaler_np(low, sd, kind: str) -> np.ndarray:
    """Independent fp64 numpy restatement of the upscaler (scatter-form transposed convs); low [1, 2, *sp]."""
    u = UPSCALERS[kind]
    k = u["keys"]
    g = lambda name: _np(sd[name])
    x = _conv_transpose_np(_np(low)[0], g(k[0] + ".weight"), g(k[0] + ".bias"), 2)
    x = 1.0 / (1.0 + np.exp(-x)) if u["act"] == "sigmoid" else np.tanh(x)
    if u["layers"] == 2:
        x = _conv_transpose_np(x, g(k[1] + ".weight"), g(k[1] + ".bias"), u["stride2"])
    w3 = g(k[-1] + ".weight").reshape(2, -1)
    out = np.tensordot(w3, x, axes=(1, 0)) + g(k[-1] + ".bias").reshape((2,) + (1,) * (x.ndim - 1))
    return out[None]


def ic_target_torch(low: torch.Tensor, kind: str, size) -> torch.Tensor:
    """The interpolated low-resolution state `get_ic_loss` compares the upscaler with: GS2D:334 bicubic to `size`;
    GS3D:328 trilinear; BUR1:465-470 periodic extension by one row/column, bicubic with align_corners, last
    row/column dropped."""
    if kind == "gs2d":
        return F.interpolate(low, tuple(size), mode="bicubic")
    if kind == "gs3d":
        return F.interpolate(low, tuple(size), mode="trilinear")
    ext = torch.cat((low, low[:, :, :, 0:1]), dim=3)
    ext = torch.cat((ext, ext[:, :, 0:1, :]), dim=2)
    return F.interpolate(ext, tuple(n + 1 for n in size), mode="bicubic", align_corners=True)[:, :, :-1, :-1]


def ic_loss_torch(low: torch.Tensor, sd: Dict[str, torch.Tensor], kind: str, size) -> torch.Tensor:
    """`get_ic_loss(model)` (GS2D:331-338): mse(upscaler(low), interpolated low)."""
    return F.mse_loss(upscaler_torch(low, sd, kind), ic_target_torch(low, kind, size))


# --------------------------------------------------------------------------------------
# Synthetic initial states of SURVEY 8d (seeded, periodic-safe) -- shared by tests and bench.
# --------------------------------------------------------------------------------------


def ic_spiral_2d(n: int = 128, dtype=torch.float64) -> torch.Tensor:
    """cfg1: u = tanh(r) cos(theta - r), v = tanh(r) sin(theta - r) on grid (i - n/2) * 0.2."""
    ax = (torch.arange(n, dtype=torch.float64) - n // 2) * 0.2
    x, y = torch.meshgrid(ax, ax, indexing="ij")
    r = torch.sqrt(x * x + y * y)
    th = torch.atan2(y, x)
    return torch.stack((torch.tanh(r) * torch.cos(th - r), torch.tanh(r) * torch.sin(th - r)))[None].to(dtype)


def ic_gs_2d(n: int = 256, seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    """cfg2: u=1, v=0 with (n//64)^2 random square patches (half-width 16) u=.5 v=.25, + 0.01 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.ones(n, n, dtype=torch.float64)
    v = torch.zeros(n, n, dtype=torch.float64)
    hw = max(1, min(16, n // 8))
    for _ in range(max(1, (n // 64)) ** 2):
        ci, cj = (int(t) for t in torch.randint(0, n, (2,), generator=g))
        ii = (torch.arange(ci - hw, ci + hw) % n)[:, None]
        jj = (torch.arange(cj - hw, cj + hw) % n)[None, :]
        u[ii, jj] = 0.5
        v[ii, jj] = 0.25
    h = torch.stack((u, v))[None]
    h = h + 0.01 * torch.randn(h.shape, generator=g, dtype=torch.float64)
    return h.to(dtype)


def ic_gs_3d(shape, seed: int = 0, dtype=torch.float32, z0: int = 0, z_total: Optional[int] = None) -> torch.Tensor:
    """cfg4/5: u=1, v=0, centre cube (half-width N/8) u=.5 v=.25, + 0.01 * noise.

    The noise is a counter-based hash of the GLOBAL cell index so that any slab [z0, z0+shape[0])
    of a z_total-deep grid reproduces the same global field for every rank count.
    """
    d, hh, w = shape
    zt = d if z_total is None else z_total
    z = torch.arange(z0, z0 + d, dtype=torch.int64)[:, None, None]
    y = torch.arange(hh, dtype=torch.int64)[None, :, None]
    x = torch.arange(w, dtype=torch.int64)[None, None, :]
    inside = ((z - zt // 2).abs() < max(1, zt // 8)) & ((y - hh // 2).abs() < max(1, hh // 8)) & \
             ((x - w // 2).abs() < max(1, w // 8))
    u = torch.where(inside, 0.5, 1.0).to(torch.float64)
    v = torch.where(inside, 0.25, 0.0).to(torch.float64)
    lin = (z * hh + y) * w + x
    out = []
    for f, base in enumerate((u, v)):
        k = (lin * 2 + f + seed * 7919) & 0x7FFFFFFF
        k = (k * 1103515245 + 12345) & 0x7FFFFFFF
        k = ((k ^ (k >> 13)) * 1664525 + 1013904223) & 0x7FFFFFFF
        k = (k ^ (k >> 16)) & 0xFFFFFF
        noise = (k.to(torch.float64) / float(1 << 24) - 0.5) * 3.4641016  # unit variance uniform
        out.append(base + 0.01 * noise)
    return torch.stack(out)[None].to(dtype)


def ic_fourier_2d(n: int = 512, seed: int = 1, modes: int = 4, dtype=torch.float32) -> torch.Tensor:
    """cfg3: smooth random Fourier field, modes <= 4, amplitude ~1."""
    g = torch.Generator().manual_seed(seed)
    ax = torch.arange(n, dtype=torch.float64) * (2 * np.pi / n)
    x, y = torch.meshgrid(ax, ax, indexing="ij")
    fields = []
    for _ in range(2):
        f = torch.zeros(n, n, dtype=torch.float64)
        for kx in range(0, modes + 1):
            for ky in range(0, modes + 1):
                if kx == 0 and ky == 0:
                    continue
                a, ph1, ph2 = torch.rand(3, generator=g, dtype=torch.float64)
                f = f + (a - 0.5) * torch.sin(kx * x + 2 * np.pi * ph1) * torch.cos(ky * y + 2 * np.pi * ph2)
        fields.append(f / f.abs().max())
    return torch.stack(fields)[None].to(dtype)
