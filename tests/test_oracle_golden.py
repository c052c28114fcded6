"""The oracle against vectors produced by the reference's own classes (tests/golden/make_golden.py).

CPU only.  Pins both restatements in oracle/percnn_oracle.py: the ATen-op-sequence one must track
the reference to rounding, the numpy one (fp64, written from the maths) to fp32/fp64 noise.
"""
import numpy as np
import pytest
import torch

from oracle import percnn_oracle as po
from tests.helpers import DLOSS_CASES, GOLDEN, GOLDEN_CASES, load_dloss, load_golden, rel_l2, rel_linf


@pytest.mark.parametrize("tag", list(GOLDEN_CASES))
def test_torch_restatement_matches_reference_rollout(tag):
    z, params, _ = load_golden(tag)
    variant = GOLDEN_CASES[tag]
    h0 = torch.from_numpy(z["h0"])
    nstep = int(z["nstep"])
    torch.set_num_threads(1)
    outs, second_last = po.rollout_torch(h0, params, variant, nstep, range(nstep))
    traj = torch.cat(outs, 0).numpy()
    tol = 1e-13 if h0.dtype == torch.float64 else 2e-6
    assert traj.shape == z["traj"].shape
    assert rel_linf(traj, z["traj"]) <= tol
    assert rel_linf(second_last.numpy(), z["second_last"]) <= tol
    # the rollout actually moves the state (guards against a fixture where nothing happens)
    assert rel_l2(z["traj"][-1], z["traj"][0]) > 1e-4


@pytest.mark.parametrize("tag", list(GOLDEN_CASES))
def test_numpy_restatement_matches_reference_step(tag):
    z, params, _ = load_golden(tag)
    variant = GOLDEN_CASES[tag]
    h = z["traj"][0:1]
    ref = z["traj"][1:2]
    got = po.cell_step_np(h, params, variant)
    tol = 1e-12 if z["h0"].dtype == np.float64 else 5e-6
    assert rel_linf(got, ref) <= tol


@pytest.mark.parametrize("tag", list(GOLDEN_CASES))
def test_numpy_adjoint_matches_reference_autograd(tag):
    """Hand-derived adjoint replayed over the recorded trajectory vs autograd grads."""
    z, params, grads = load_golden(tag)
    variant = GOLDEN_CASES[tag]
    traj = z["traj"].astype(np.float64)
    wts = z["loss_weights"].astype(np.float64)
    nstep = int(z["nstep"])
    G = wts[nstep:nstep + 1].copy()
    acc = {}
    for t in range(nstep - 1, -1, -1):
        G, pg = po.cell_step_vjp_np(traj[t:t + 1], G, params, variant)
        G = G + wts[t:t + 1]
        for k, v in pg.items():
            acc[k] = acc.get(k, 0) + v
    tol = 1e-10 if z["h0"].dtype == np.float64 else 3e-4
    assert rel_l2(G, z["g_h0"]) <= tol
    assert set(grads) <= set(acc), (set(grads) - set(acc))
    for k, ref in grads.items():
        assert rel_l2(np.asarray(acc[k]).reshape(ref.shape), ref) <= tol, k


def test_torch_autograd_of_oracle_matches_reference_grads():
    z, params, grads = load_golden("gs2d")
    p = {k: v.clone().requires_grad_(k in grads) for k, v in params.items()}
    h0 = torch.from_numpy(z["h0"]).requires_grad_(True)
    nstep = int(z["nstep"])
    outs, _ = po.rollout_torch(h0, p, "gs2d", nstep, range(nstep))
    loss = (torch.cat(outs, 0) * torch.from_numpy(z["loss_weights"])).sum()
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    assert rel_l2(h0.grad.numpy(), z["g_h0"]) <= 1e-5
    for k, ref in grads.items():
        assert rel_l2(p[k].grad.numpy(), ref) <= 2e-4, k


@pytest.mark.parametrize("tag", list(DLOSS_CASES))
def test_data_loss_restatement_matches_reference(tag):
    """Frame selection, loss value and gradient of the scripts' data loss (GS3D:394-403 and siblings) against the
    reference's own loop + nn.MSELoss + autograd; the gradient goes through the hand-derived adjoint with the loss
    gradient injected at the selected states -- the scheme the fused kernels implement."""
    z, params, grads = load_dloss(tag)
    variant = DLOSS_CASES[tag]
    nstep, ts, ss, ff = int(z["nstep"]), int(z["t_stride"]), int(z["s_stride"]), int(z["first_frames"])
    frames = po.data_loss_frames(nstep, range(nstep), ts, None if ff < 0 else ff)
    assert frames == [int(f) for f in z["frames"]]
    traj = z["traj"].astype(np.float64)
    fp64 = z["h0"].dtype == np.float64
    loss = po.data_loss_np(traj, z["truth_sub"], frames, ss)
    assert abs(loss - float(z["loss"])) <= (1e-12 if fp64 else 2e-6) * abs(float(z["loss"]))
    gl = po.data_loss_grad_np(traj, z["truth_sub"], frames, ss, float(z["gscale"]))
    assert np.count_nonzero(gl[nstep]) == 0          # `[0:-1:...]` never selects the last (dummy) state
    G = gl[nstep:nstep + 1].copy()
    acc = {}
    for t in range(nstep - 1, -1, -1):
        G, pg = po.cell_step_vjp_np(traj[t:t + 1], G, params, variant)
        G = G + gl[t:t + 1]
        for k, v in pg.items():
            acc[k] = acc.get(k, 0) + v
    tol = 1e-10 if fp64 else 3e-4
    assert rel_l2(G, z["g_h0"]) <= tol
    for k, ref in grads.items():
        assert rel_l2(np.asarray(acc[k]).reshape(ref.shape), ref) <= tol, k


def test_data_loss_frames_follow_effective_step():
    """`outputs` only holds the effective steps, so list index != state index in general (GS3D:191-212)."""
    assert po.data_loss_frames(7, [0, 2, 3, 6], 2) == [0, 3]          # outputs = h0 h1 h3 h4 h7 -> [0:-1:2] = h0 h3
    assert po.data_loss_frames(6, range(6), 5) == [0, 5]
    assert po.data_loss_frames(30, range(30), 15) == [0, 15]
    assert po.data_loss_frames(31, range(31), 15) == [0, 15, 30]
    assert po.data_loss_frames(9, range(9), 4, first_frames=2) == [0, 4]


@pytest.mark.parametrize("tag", ["fwd", "gs2d", "gs3d"])
def test_physics_loss_restatements_match_reference(tag):
    """Both restatements of `loss_gen(output, loss_generator())` (FWD:288-357, GS2D:270-353, GS3D:286-345) against
    the reference's own loss module run on the reference's own trajectory: value and dloss/doutput."""
    import os
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, f"phys_{tag}.npz"))
    fp64 = z["h0"].dtype == np.float64
    out = torch.from_numpy(z["traj"]).requires_grad_(True)
    loss = po.phys_loss_torch(out, tag)
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) <= (1e-12 if fp64 else 2e-5) * abs(float(z["loss"]))
    assert rel_l2(out.grad.numpy(), z["g_traj"]) <= (1e-11 if fp64 else 1e-4)
    # numpy version in fp64 on the periodic grid with the double-counting weights (measured: 1e-7 / 2e-7 against
    # the fp32 reference, 1e-14 against the fp64 one)
    l64, g64 = po.phys_loss_np(z["traj"], tag)
    assert abs(l64 - float(z["loss"])) <= (1e-12 if fp64 else 2e-6) * abs(float(z["loss"]))
    assert rel_l2(g64, z["g_traj"]) <= (1e-11 if fp64 else 2e-6)
    assert np.count_nonzero(z["g_traj"][-1]) == 0        # the last frame never enters (FWD:318-327)


def test_stencil_tables_match_reference_weights():
    z, params, _ = load_golden("gs3d")
    w = params["W_laplace.weight"].numpy()
    np.testing.assert_allclose(w, po.laplace_stencil(3) / (100 / 48) ** 2, rtol=1e-6)
    assert np.count_nonzero(w) == 13
    z, params, _ = load_golden("bur3")
    np.testing.assert_array_equal(params["dx_op.filter.weight"].numpy(), po.dx_stencil_2d())
    np.testing.assert_array_equal(params["dy_op.filter.weight"].numpy(), po.dy_stencil_2d())
    np.testing.assert_array_equal(params["laplace_op.filter.weight"].numpy(), po.laplace_stencil(2))


def test_fwd_checkpoint_known_answer():
    """with rcnn_pde.pt the Pi-block reproduces the analytic lambda-omega reaction
    (1-A)u + A v, -A u + (1-A) v with A = u^2+v^2 (to ~1e-7 max-abs in fp64)."""
    from tests.helpers import load_weights
    p = load_weights("fwd")
    g = np.random.default_rng(0)
    u = g.uniform(-1, 1, (16, 16))
    v = g.uniform(-1, 1, (16, 16))
    h = torch.tensor(np.stack((u, v))[None])
    got = po.cell_step_np(h, p, "fwd")
    lap = lambda a: po._apply_taps_np(a, po.laplace_stencil(2)[0, 0] / 0.2 ** 2)
    A = u * u + v * v
    fu = float(p["DA"]) * lap(u) + (1 - A) * u + A * v
    fv = float(p["DB"]) * lap(v) - A * u + (1 - A) * v
    want = np.stack((u + 0.0125 * fu, v + 0.0125 * fv))[None]
    assert np.abs(got - want).max() < 1e-6


@pytest.mark.parametrize("tag,variant", [("bur3", "bur3"), ("lo3", "lo3"), ("lo3n", "lo3")])
def test_oracle_rk4_matches_reference_forward_rk4(tag, variant):
    """tests/golden/rk4_*.npz: the reference's own `forward_rk4` (BUR3:159-206) for three steps."""
    import os
    z = np.load(os.path.join(GOLDEN, f"rk4_{tag}.npz"))
    p = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    h = torch.from_numpy(z["h0"])
    for t in range(3):
        h = po.rk4_step_torch(h, p, variant)
        assert rel_l2(h.numpy(), z["traj"][t + 1:t + 2]) <= 1e-13, t
