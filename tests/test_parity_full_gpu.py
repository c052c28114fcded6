"""GPU parity at BASELINE.json's OWN config sizes and rollout lengths.

Expected values come from the reference's own `RCNNCell` classes run at full size in the build container
(tests/golden/make_golden_full.py -> tests/golden/full_*.npz): cfg1 128^2 fp64 x 200 steps, cfg2 256^2 x 1000
steps, cfg3 512^2 40-step back-propagation (V-BUR1 fp32 and V-BUR3 fp64), cfg4 128^3 x 500 steps, cfg5's grid
(512^3) x 3 steps.  fp32 cases also carry the same class evaluated in fp64 and the reference's own fp32-vs-fp64
distance, so each test checks BOTH

    rel-L2(new, reference fp32) <= 1e-5                                  (north_star tolerance), and
    rel-L2(new, fp64 yardstick) <= max(2 x rel-L2(reference fp32, fp64), 1e-6)   (SURVEY 8c).

Both evaluation orders of the 1x1 Pi-block are covered: the folded bivariate cubic (default) and the reference's
channel-by-channel order (PERCNN_FLAG_EVAL_BRANCH).  Initial states are regenerated from their seeds
(oracle.percnn_oracle.ic_*) and verified against the stored checksum.
"""
import os

import numpy as np
import pytest
import torch

from oracle import percnn_oracle as po
from percnn_b200 import _lib, engine
from tests.helpers import GOLDEN, checksums_match, load_weights, make_cell, rel_l2, rel_linf, state_checksum

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _full(name):
    return np.load(os.path.join(GOLDEN, f"full_{name}.npz"))


def _check_ic(h0, z, key="h0_checksum"):
    assert checksums_match(state_checksum(h0.cpu()), z[key]), "seeded initial state differs from the one the golden was made from"


def _cell(tag, weights=None, flags=0):
    cell = make_cell(tag)
    if weights is not None:
        cell.load_state_dict(weights, strict=True)
    cell._flags = flags
    return cell.to(DEV)


def _yardstick_ok(new, z, key_new_vs_ref, ref32_vs_f64, new_vs_f64):
    assert key_new_vs_ref <= 1e-5, f"vs reference fp32: {key_new_vs_ref:.3e}"
    bound = max(2.0 * float(ref32_vs_f64), 1e-6)
    assert new_vs_f64 <= bound, f"vs fp64 yardstick: {new_vs_f64:.3e} > {bound:.3e} (reference's own fp32-vs-fp64: {float(ref32_vs_f64):.3e})"


# ---- cfg1 ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", [0, _lib.FLAG_EVAL_BRANCH], ids=["folded", "branch"])
def test_cfg1_lambda_omega_128_fp64_200_steps(flags):
    z = _full("cfg1")
    h0 = po.ic_spiral_2d(128)
    _check_ic(h0, z)
    cell = _cell("fwd", load_weights("fwd"), flags)
    with torch.no_grad():
        states = cell.rollout(h0.to(DEV), 200)
        # final-state-only path (persistent multi-step kernel / ping-pong) must agree with the taped rollout
        _, fin = cell.rollout_emit(h0.to(DEV), 200, [False] * 200, want_final=True)
    for t in (1, 50, 200):
        err = rel_l2(states[t].cpu().numpy(), z[f"state_{t}"][0])
        assert err <= 1e-11, (t, err)
    assert rel_linf(states[200].cpu().numpy(), z["state_200"][0]) <= 1e-10
    assert torch.equal(fin, states[200])


# ---- cfg2 ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", [0, _lib.FLAG_EVAL_BRANCH], ids=["folded", "branch"])
def test_cfg2_gray_scott_256_1000_steps(flags):
    z = _full("cfg2")
    h0 = po.ic_gs_2d(256, seed=0)
    _check_ic(h0, z)
    cell = _cell("gs2d", load_weights("gs2d"), flags)
    with torch.no_grad():
        states = cell.rollout(h0.to(DEV), 1000)
        _, fin = cell.rollout_emit(h0.to(DEV), 1000, [False] * 1000, want_final=True)
    s1, s100, s1000 = (states[t].cpu().numpy() for t in (1, 100, 1000))
    assert rel_linf(s1, z["state_1"][0]) <= 1e-6
    assert rel_l2(s100[..., ::2, ::2], z["state_100"][0]) <= 1e-5
    _yardstick_ok(s100, z, rel_l2(s100[..., ::2, ::2], z["state_100"][0]), z["ref32_vs_f64_l2"][1],
                  rel_l2(s100[..., ::2, ::2], z["f64_state_100"][0]))
    _yardstick_ok(s1000, z, rel_l2(s1000, z["state_1000"][0]), z["ref32_vs_f64_l2"][2],
                  rel_l2(s1000[..., ::2, ::2], z["f64_state_1000"][0]))
    # L-inf at 1000 steps: the reference's own floor is 5e-6 (fp32 vs fp64)
    assert rel_linf(s1000, z["state_1000"][0]) <= max(4 * float(z["ref32_vs_f64_linf"][2]), 1e-5)
    assert torch.equal(fin, states[1000])


# ---- cfg3 ------------------------------------------------------------------------------------------------------
def _cfg3_inputs(z, dtype):
    h0 = po.ic_fourier_2d(512, seed=1, dtype=dtype)
    tgt = po.ic_fourier_2d(512, seed=2, dtype=dtype)[:, :, ::2, ::2].expand(8, -1, -1, -1).contiguous()
    _check_ic(h0, z)
    _check_ic(tgt[0], z, "target_checksum")
    return h0, tgt


def _bptt_dense(cell, h0, tgt):
    """The scripts' way: loss on slices of the concatenated trajectory, autograd through the fused rollout."""
    hd = h0.to(DEV).requires_grad_(True)
    states = cell.rollout(hd, 40)
    loss = torch.mean((states[0:-1:5, :, ::2, ::2] - tgt.to(DEV)) ** 2)           # BUR1:610-614
    loss.backward(retain_graph=True)
    return loss, hd.grad, states[-1].detach()


def _bptt_fused(cell, h0, tgt):
    """Same loss through the fused data-loss kernels (gradient injected inside the adjoint)."""
    hd = h0.to(DEV).requires_grad_(True)
    sel = [(t % 5 == 0) and t < 40 for t in range(41)]
    states, loss = cell.rollout_data_loss(hd, 40, tgt.to(DEV), sel, 2)
    loss.backward()
    return loss, hd.grad, states[-1].detach()


@pytest.mark.parametrize("mode", ["dense", "fused"])
def test_cfg3_burgers_512_pi_block_bptt_40_steps(mode):
    """V-BUR1 (5x5 Pi-block, hc=16, fp32, shipped Stage-1 checkpoint): loss, dL/dh0 and all 18 parameter gradients."""
    z = _full("cfg3_bur1")
    h0, tgt = _cfg3_inputs(z, torch.float32)
    cell = _cell("bur1", load_weights("bur1"))
    loss, g_h0, fin = (_bptt_dense if mode == "dense" else _bptt_fused)(cell, h0, tgt)
    assert abs(loss.item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    assert abs(loss.item() - float(z["f64_loss"])) <= 1e-6 * abs(float(z["f64_loss"]))
    g = g_h0.cpu().numpy()[..., ::4, ::4]
    _yardstick_ok(g, z, rel_l2(g, z["g_h0_sub"]), z["ref32_vs_f64_g_h0"], rel_l2(g, z["f64_g_h0_sub"]))
    f = fin.cpu().numpy()[None][..., ::4, ::4]
    _yardstick_ok(f, z, rel_l2(f, z["final_sub"]), z["ref32_vs_f64_final"], rel_l2(f, z["f64_final_sub"]))
    named = dict(cell.named_parameters())
    for key in [k for k in z.files if k.startswith("grad/")]:
        name = key[len("grad/"):]
        got = named[name].grad.cpu().numpy()
        floor = float(z["ref32_vs_f64_grad/" + name])       # the reference's own fp32 noise on this tensor (up to 2.5e-5)
        e64 = rel_l2(got, z["f64_grad/" + name])
        e32 = rel_l2(got, z[key])
        assert e64 <= max(2.0 * floor, 2e-6), (name, e64, floor)
        assert e32 <= max(1e-5, 3.0 * floor), (name, e32, floor)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-5)], ids=["fp64", "fp32"])
@pytest.mark.parametrize("mode", ["dense", "fused"])
def test_cfg3_burgers_512_advection_stencil_bptt_40_steps(dtype, tol, mode):
    """V-BUR3 (d/dx, d/dy advection stencils, script literals): fp64 like the reference, and the fp32 variant
    against the same fp64 vectors at north_star's 1e-5."""
    z = _full("cfg3_bur3")
    h0, tgt = _cfg3_inputs(z, torch.float64)
    cell = make_cell("bur3")
    if dtype == torch.float32:
        cell = cell.float()
        cell.dtype = torch.float32
    cell = cell.to(DEV)
    loss, g_h0, fin = (_bptt_dense if mode == "dense" else _bptt_fused)(cell, h0.to(dtype), tgt.to(dtype))
    assert abs(loss.item() - float(z["loss"])) <= tol * abs(float(z["loss"]))
    assert rel_l2(g_h0.cpu().numpy()[..., ::4, ::4], z["g_h0_sub"]) <= tol
    assert rel_l2(fin.cpu().numpy()[None][..., ::4, ::4], z["final_sub"]) <= tol
    named = dict(cell.named_parameters())
    for key in [k for k in z.files if k.startswith("grad/")]:
        got = named[key[len("grad/"):]].grad.cpu().numpy()
        assert rel_l2(got, z[key]) <= (tol if dtype == torch.float64 else 2e-4), key   # fp32 scalar sums over 1e7 terms


# ---- cfg4 / cfg5 -----------------------------------------------------------------------------------------------
def _ic_gs3d_on_gpu(shape):
    """oracle.ic_gs_3d evaluated on the GPU (integer hash + IEEE fp64 arithmetic: identical to the CPU result)."""
    with torch.device(DEV):
        return po.ic_gs_3d(shape, seed=0)


@pytest.mark.parametrize("flags", [0, _lib.FLAG_EVAL_BRANCH], ids=["tma_folded", "generic_branch"])
def test_cfg4_gray_scott_128_cubed_500_steps(flags):
    z = _full("cfg4")
    h0 = _ic_gs3d_on_gpu((128, 128, 128))
    _check_ic(h0, z)
    cell = _cell("gs3d", load_weights("gs3d"), flags)
    assert engine.get_plan(cell._spec(), (128, 128, 128), torch.device(DEV)).uses_tma == (flags == 0)
    emit = [s + 1 in (1, 50, 500) for s in range(500)]
    with torch.no_grad():
        traj, fin = cell.rollout_emit(h0, 500, emit, want_final=True)
    assert torch.equal(traj[2], fin)
    for i, t in enumerate((1, 50, 500)):
        s = traj[i].cpu().numpy()[None]
        e_plane = rel_l2(s[:, :, 64], z[f"state_{t}_plane64"])
        e_sub = rel_l2(s[..., ::4, ::4, ::4], z[f"state_{t}_sub"])
        e_f64 = rel_l2(s[..., ::4, ::4, ::4], z[f"f64_state_{t}_sub"])
        assert e_plane <= 1e-5, (t, e_plane)
        _yardstick_ok(s, z, e_sub, z["ref32_vs_f64_l2"][i], e_f64)
        assert abs(float(np.linalg.norm(s.astype(np.float64))) - float(z[f"norm_{t}"])) <= 1e-6 * float(z[f"norm_{t}"])
    assert rel_linf(traj[0].cpu().numpy()[None][..., ::4, ::4, ::4], z["state_1_sub"]) <= 1e-6


def test_cfg5_grid_512_cubed_first_steps_match_the_reference():
    """cfg5's grid: the reference can only afford 3 forward steps at 512^3 on the build container; they pin the TMA
    kernel at full width (4 x 37 tiles, every seam) against the reference's own cell."""
    z = _full("cfg5")
    h0 = _ic_gs3d_on_gpu((512, 512, 512))
    _check_ic(h0, z)
    cell = _cell("gs3d", load_weights("gs3d"))
    assert engine.get_plan(cell._spec(), (512, 512, 512), torch.device(DEV)).uses_tma
    with torch.no_grad():
        traj, _ = cell.rollout_emit(h0, 3, [True, False, True])
    for i, t in enumerate((1, 3)):
        s = traj[i][None]
        assert rel_linf(s[..., ::16, ::16, ::16].cpu().numpy(), z[f"state_{t}_sub"]) <= 2e-6
        assert rel_linf(s[:, :, 0, ::2, ::2].cpu().numpy(), z[f"state_{t}_plane0"]) <= 2e-6
        assert rel_linf(s[:, :, ::4, 255, :].cpu().numpy(), z[f"state_{t}_row255"]) <= 2e-6
        nrm = float(s.double().norm())
        assert abs(nrm - float(z[f"norm_{t}"])) <= 1e-6 * float(z[f"norm_{t}"])
