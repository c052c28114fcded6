"""GPU parity tests of the fused data loss (include/percnn_b200.h percnn_data_loss_t).

The CUDA path -- `percnn_data_loss_fwd` for the value, the loss gradient injected inside the adjoint kernels
(`percnn_rollout_bwd_loss`) -- against
 (1) vectors recorded from the reference's own loop + nn.MSELoss + autograd (tests/golden/dloss_*.npz),
 (2) the library's own dense path (slice the states with stock torch ops, dense gradient tape) on ragged sizes
     and strides, where both must agree to rounding,
 (3) the numpy oracle (oracle.percnn_oracle.data_loss_np / data_loss_grad_np).

Tolerances: loss 1e-5 relative (fp32) / 1e-12 (fp64); gradients rel-L2 1e-5 (fp32; 2e-4 for the scalar diffusion
coefficients, see tests/test_parity_gpu.py) / 1e-10 (fp64).
"""
import numpy as np
import pytest
import torch

from oracle import percnn_oracle as po
from percnn_b200 import _lib, engine
from tests.helpers import DLOSS_CASES, load_dloss, load_golden, make_cell, rel_l2

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _cell(tag, params):
    cell = make_cell(tag)
    cell.load_state_dict(params, strict=True)
    return cell.to(DEV)


def _sel(nstep, frames):
    sel = [False] * (nstep + 1)
    for f in frames:
        sel[int(f)] = True
    return sel


def _sub(nd, s):
    return (slice(None), slice(None)) + (slice(None, None, s),) * nd


@pytest.mark.parametrize("tag", list(DLOSS_CASES))
def test_fused_data_loss_matches_reference_golden(tag):
    z, params, grads = load_dloss(tag)
    cell = _cell(tag, params)
    fp64 = z["h0"].dtype == np.float64
    nstep, ss = int(z["nstep"]), int(z["s_stride"])
    h0 = torch.from_numpy(z["h0"]).to(DEV).requires_grad_(True)
    truth = torch.from_numpy(z["truth_sub"]).to(DEV)
    states, loss = cell.rollout_data_loss(h0, nstep, truth, _sel(nstep, z["frames"]), ss)
    assert loss.dim() == 0 and loss.dtype == h0.dtype
    assert abs(loss.item() - float(z["loss"])) <= (1e-12 if fp64 else 1e-5) * abs(float(z["loss"]))
    (float(z["gscale"]) * loss).backward(retain_graph=True)
    gtol = 1e-10 if fp64 else 1e-5
    assert rel_l2(h0.grad.cpu().numpy(), z["g_h0"]) <= gtol
    named = dict(cell.named_parameters())
    for k, ref in grads.items():
        got = named[k].grad
        assert got is not None, k
        tol = gtol if ref.ndim > 0 or fp64 else 2e-4
        assert rel_l2(got.cpu().numpy(), ref) <= tol, (k, rel_l2(got.cpu().numpy(), ref))
    # the states returned alongside are the ordinary rollout
    assert rel_l2(states.detach().cpu().numpy(), z["traj"]) <= (1e-11 if fp64 else 1e-5)


CASES = [  # tag, spatial shape, nsteps, selected states, stride  (ragged extents, strides that do not divide them)
    ("gs2d", (37, 50), 7, (0, 3, 6), 3),
    ("gs2d", (32, 32), 5, (1, 5), 1),           # stride 1 and the LAST state selected
    ("fwd", (24, 28), 6, (0, 2, 4), 5),
    ("gs3d", (7, 9, 20), 5, (0, 2, 4), 2),      # generic 3-D kernel
    ("gs3d", (8, 30, 128), 5, (0, 2, 5), 2),    # TMA kernel, shifted-back last tile, last state selected
    ("gs3d", (9, 16, 256), 4, (1, 3), 3),       # TMA kernel, stride 3 (x lattice not aligned with the lanes' quads)
    ("gs3d", (8, 16, 128), 4, (0, 3), 4),       # TMA kernel, stride 4
    ("bur1", (20, 36), 4, (0, 2, 4), 2),        # 5x5 Pi-block (separate injection pass)
    ("lo1", (16, 32), 3, (0, 3), 3),
    ("bur3", (21, 24), 5, (0, 5), 2),
    ("lo3", (20, 23), 5, (1, 4), 2),
    ("lo3n", (20, 24), 4, (0, 2), 4),
]


@pytest.mark.parametrize("tag,shape,nstep,frames,stride", CASES)
def test_fused_equals_dense_path_and_oracle(tag, shape, nstep, frames, stride):
    """Same cell, same inputs: (a) fused loss + injected gradient, (b) stock torch slicing of the states + dense
    gradient tape through the same adjoint kernels, (c) numpy oracle for the value and for dL/dstates."""
    _, params, _ = load_golden(tag)
    res = {}
    g = torch.Generator().manual_seed(3)
    nd = len(shape)
    dtype = make_cell(tag).dtype
    h0_cpu = (0.2 + 0.6 * torch.rand((1, 2, *shape), generator=g, dtype=torch.float64)).to(dtype)
    if tag in ("fwd", "lo1", "lo3", "lo3n", "bur1", "bur3"):
        h0_cpu = h0_cpu - 0.5
    low = tuple((n + stride - 1) // stride for n in shape)
    truth_cpu = torch.rand((len(frames), 2, *low), generator=g, dtype=torch.float64).to(dtype)
    for mode in ("fused", "dense"):
        cell = _cell(tag, params)
        h0 = h0_cpu.to(DEV).requires_grad_(True)
        truth = truth_cpu.to(DEV)
        if mode == "fused":
            states, loss = cell.rollout_data_loss(h0, nstep, truth, _sel(nstep, frames), stride)
        else:
            states = cell.rollout(h0, nstep)
            loss = torch.nn.MSELoss()(states[list(frames)][_sub(nd, stride)], truth)
        (2.5 * loss).backward()
        res[mode] = (loss.item(), h0.grad.cpu().numpy(), states.detach().cpu().numpy(),
                     {k: p.grad.cpu().numpy() for k, p in cell.named_parameters() if p.grad is not None})
    fp64 = dtype == torch.float64
    tol = 1e-12 if fp64 else 2e-6
    assert abs(res["fused"][0] - res["dense"][0]) <= (1e-13 if fp64 else 2e-6) * abs(res["dense"][0])
    assert np.array_equal(res["fused"][2], res["dense"][2])
    assert rel_l2(res["fused"][1], res["dense"][1]) <= tol
    assert set(res["fused"][3]) == set(res["dense"][3])
    for k in res["dense"][3]:
        assert rel_l2(res["fused"][3][k], res["dense"][3][k]) <= max(tol, 1e-5 if not fp64 else tol), k
    # oracle: value, and dL/dh0 for a zero-step-free check of the injected lattice via the numpy gradient of the loss
    want = po.data_loss_np(res["fused"][2], truth_cpu.numpy(), list(frames), stride)
    assert abs(res["fused"][0] - want) <= (1e-12 if fp64 else 1e-5) * abs(want)


def test_injected_gradient_lattice_single_step():
    """One adjoint step with a zero incoming gradient returns exactly the injected loss gradient: compare with the
    numpy oracle's dL/dstate cell by cell (TMA plan and generic plan, stride 2 and 3)."""
    _, params, _ = load_golden("gs3d")
    for shape, stride, flags in (((8, 16, 128), 2, 0), ((8, 18, 128), 3, 0), ((8, 16, 128), 2, _lib.FLAG_NO_TMA),
                                 ((5, 7, 11), 2, 0)):
        cell = _cell("gs3d", params)
        cell._flags = flags
        g = torch.Generator().manual_seed(11)
        h = torch.rand((1, 2, *shape), generator=g)
        low = tuple((n + stride - 1) // stride for n in shape)
        target = torch.rand((1, 2, *low), generator=g)
        plan = cell._plan(h.to(DEV))
        assert plan.uses_tma == (flags == 0 and shape[2] % 128 == 0)
        flat = engine.pack_params(cell._packed_tensors(), torch.float32)
        plan.params_load(flat)
        hd = h[0].to(DEV).contiguous()
        gout = torch.zeros_like(hd)
        gin = torch.empty_like(hd)
        gscale = torch.tensor([0.75], device=DEV)
        n_total = 2 * int(np.prod(low))
        plan.param_grads_begin()
        plan.step_bwd_loss(hd, gout, gin, target_frame=target[0].to(DEV).contiguous(), stride=stride, n_total=n_total,
                           gscale=gscale)
        want = po.data_loss_grad_np(h.numpy(), target.numpy(), [0], stride, 0.75)[0]
        got = gin.cpu().numpy()
        assert np.count_nonzero(got) == np.count_nonzero(want)
        assert rel_l2(got, want) <= 1e-6


def test_loss_and_dense_gradient_combine():
    """A loss that uses BOTH the fused data loss and the states themselves (e.g. a physics loss on `output`)."""
    z, params, _ = load_dloss("gs2d")
    nstep, ss = int(z["nstep"]), int(z["s_stride"])
    truth = torch.from_numpy(z["truth_sub"]).to(DEV)
    w = torch.rand((nstep + 1, *z["h0"].shape[1:]), generator=torch.Generator().manual_seed(1)).to(DEV)
    out = []
    for mode in ("fused", "dense"):
        cell = _cell("gs2d", params)
        h0 = torch.from_numpy(z["h0"]).to(DEV).requires_grad_(True)
        if mode == "fused":
            states, ld = cell.rollout_data_loss(h0, nstep, truth, _sel(nstep, z["frames"]), ss)
        else:
            states = cell.rollout(h0, nstep)
            ld = torch.nn.MSELoss()(states[[int(f) for f in z["frames"]]][_sub(2, ss)], truth)
        (10 * ld + 1e-3 * (states * w).sum()).backward()
        out.append((h0.grad.cpu().numpy(), cell.CA.grad.item(), cell.Wh1_u.weight.grad.cpu().numpy()))
    assert rel_l2(out[0][0], out[1][0]) <= 2e-6
    assert abs(out[0][1] - out[1][1]) <= 1e-5 * abs(out[1][1])
    assert rel_l2(out[0][2], out[1][2]) <= 1e-5


def test_rcnn_forward_data_loss_drop_in():
    """FusedRCNN.forward_data_loss == the script's lines GS2D:393-401 on the module's own forward()."""
    from percnn_b200.variants import gs2d
    z = np.load(__import__("os").path.join(__import__("tests.helpers", fromlist=["GOLDEN"]).GOLDEN, "rcnn_gs2d.npz"))
    low = torch.from_numpy(z["init_state_low"]).to(DEV)
    step, eff = 9, [0, 1, 3, 4, 5, 7, 8]
    res = []
    for mode in ("fused", "script"):
        model = gs2d.RCNN(input_channels=2, hidden_channels=8, init_state_low=low, input_kernel_size=5, step=step,
                          effective_step=eff).to(DEV)
        sd = {k[len("state/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state/")}
        model.load_state_dict(sd, strict=True)
        frames = po.data_loss_frames(step, eff, 2)
        hw = model.UpconvBlock(low).shape[2:]
        truth = torch.rand((len(frames), 2, (hw[0] + 2) // 3, (hw[1] + 2) // 3),
                           generator=torch.Generator().manual_seed(2)).to(DEV)
        if mode == "fused":
            outputs, second_last, loss = model.forward_data_loss(truth, 2, 3)
        else:
            outputs, second_last = model()
            output = torch.cat(tuple(outputs), dim=0)
            loss = torch.nn.MSELoss()(output[0:-1:2, :, ::3, ::3], truth)
        loss.backward()
        res.append((loss.item(), torch.cat(tuple(outputs), 0).detach().cpu().numpy(), second_last.detach().cpu().numpy(),
                    {k: p.grad.cpu().numpy() for k, p in model.named_parameters() if p.grad is not None}))
    assert abs(res[0][0] - res[1][0]) <= 2e-6 * abs(res[1][0])
    # (two model instances: the stock cuDNN transposed convs of the upscaler are not bit-reproducible across calls)
    assert rel_l2(res[0][1], res[1][1]) <= 1e-6 and rel_l2(res[0][2], res[1][2]) <= 1e-6
    assert set(res[0][3]) == set(res[1][3]) and any(k.startswith("UpconvBlock") for k in res[0][3])
    for k in res[1][3]:
        assert rel_l2(res[0][3][k], res[1][3][k]) <= 2e-5, k


def test_data_loss_argument_errors():
    _, params, _ = load_golden("gs2d")
    cell = _cell("gs2d", params)
    h0 = torch.rand((1, 2, 16, 16), device=DEV)
    with pytest.raises(ValueError):      # wrong target shape
        cell.rollout_data_loss(h0, 3, torch.zeros((1, 2, 5, 5), device=DEV), [True, False, False, False], 2)
    with pytest.raises(ValueError):      # mask length
        cell.rollout_data_loss(h0, 3, torch.zeros((1, 2, 8, 8), device=DEV), [True, False, False], 2)
    with pytest.raises(RuntimeError):    # CPU target: no fallback
        cell.rollout_data_loss(h0, 3, torch.zeros((1, 2, 8, 8)), [True, False, False, False], 2)
    with pytest.raises(_lib.PercnnError):  # nothing selected
        cell.rollout_data_loss(h0, 3, torch.zeros((0, 2, 8, 8), device=DEV), [False] * 4, 2)
