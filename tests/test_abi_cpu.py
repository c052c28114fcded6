"""CPU-only checks of the boundary: the C-ABI library loads and exports what the header declares, the
drop-in modules expose the reference's parameter names/shapes, and nothing silently falls back to CPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from percnn_b200 import _lib, engine
from tests.helpers import GOLDEN_CASES, load_golden, make_cell

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "percnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(percnn_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _header_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(_lib.EXPORTS) == declared
    for name in declared:
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (percnn_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    assert L.percnn_abi_version() == _lib.ABI_VERSION


def test_library_does_not_link_torch_or_the_oracle():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_plan_create_fails_loudly_without_a_gpu():
    L = _lib.lib()
    assert L.percnn_device_ok(0) == 0
    d = _lib.Desc()
    d.abi_version, d.ndim, d.dtype, d.cell, d.ksize, d.hidden = _lib.ABI_VERSION, 2, 0, 0, 1, 4
    d.extent[0], d.extent[1], d.extent[2] = 1, 8, 8
    d.dt, d.dx, d.mu_up = 0.1, 0.1, 1.0
    h = ctypes.c_void_p()
    rc = L.percnn_plan_create(ctypes.byref(d), ctypes.byref(h))
    assert rc == 4 and not h.value
    assert b"sm_100" in L.percnn_last_error()


def test_descriptor_validation_messages():
    L = _lib.lib()
    d = _lib.Desc()
    h = ctypes.c_void_p()
    d.abi_version = 99
    assert L.percnn_plan_create(ctypes.byref(d), ctypes.byref(h)) == 1
    assert b"abi_version" in L.percnn_last_error()
    d.abi_version, d.ndim = _lib.ABI_VERSION, 4
    assert L.percnn_plan_create(ctypes.byref(d), ctypes.byref(h)) == 1
    assert L.percnn_plan_create(None, ctypes.byref(h)) == 1
    assert L.percnn_param_count(None) == -1


@pytest.mark.parametrize("tag", list(GOLDEN_CASES))
def test_dropin_cell_has_reference_parameter_names_and_shapes(tag):
    """strict load_state_dict of the parameters recorded from the reference class."""
    _, params, grads = load_golden(tag)
    cell = make_cell(tag)
    sd = cell.state_dict()
    assert list(sd.keys()) == list(params.keys())
    for k, v in params.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
        assert sd[k].dtype == v.dtype, k
    cell.load_state_dict(params, strict=True)
    trainable = {n for n, p in cell.named_parameters() if p.requires_grad}
    assert trainable == set(grads.keys())


def test_constructor_side_effects_match_reference():
    """np.random.seed(1234) inside the constructor and the CA/CB draw (GS2D:60-62, BUR1:98-100)."""
    from percnn_b200.variants import burgers_stage1, gs2d
    c = gs2d.RCNNCell(2, 8, 5)
    np.random.seed(1234)
    ca, cb = (np.random.rand() - 0.5) * 2, (np.random.rand() - 0.5) * 2
    assert c.CA.item() == pytest.approx(np.float32(ca)) and c.CB.item() == pytest.approx(np.float32(cb))
    assert c.input_kernel_size == 5 and c.mu_up == 3.99e-5 and c.dt == 0.5 and c.dx == 0.01
    assert all(float(f.bias.abs().max()) == 0.0 for f in c.filter_list) and len(c.filter_list) == 8
    b = burgers_stage1.RCNNCell(2, 4, 2, 5, 1, 2)
    np.random.seed(1234)
    assert b.CA.item() == pytest.approx(np.float32(np.random.rand()))
    assert b.Wh1_u.weight.shape == (16, 2, 5, 5) and b.nu_up == 0.01


def test_laplacian_tables_match_golden():
    from percnn_b200.cells import derivative_table, laplace_table
    _, p3, _ = load_golden("gs3d")
    np.testing.assert_allclose(p3["W_laplace.weight"].numpy(), laplace_table(3) / (100 / 48) ** 2, rtol=1e-6)
    _, pb, _ = load_golden("bur3")
    np.testing.assert_array_equal(pb["dx_op.filter.weight"].numpy(), derivative_table(0))
    np.testing.assert_array_equal(pb["dy_op.filter.weight"].numpy(), derivative_table(1))
    np.testing.assert_array_equal(pb["laplace_op.filter.weight"].numpy(), laplace_table(2))


def test_cpu_tensors_are_rejected_not_silently_computed():
    cell = make_cell("gs2d")
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        cell(torch.zeros(1, 2, 8, 8))


def test_fwd_checkpoint_prefix_shim():
    from percnn_b200.variants import lambda_omega_fwd
    from tests.helpers import load_weights
    w = load_weights("fwd")
    m = lambda_omega_fwd.RCNN(1, torch.zeros(1, 2, 8, 8), 1, 0, step=2, effective_step=[0, 1])
    m.load_state_dict({"crnn_cell." + k: v for k, v in w.items()}, strict=True)
    assert torch.equal(m.rcnn_cell.DA.detach(), w["DA"])


def test_pack_params_is_state_dict_order():
    cell = make_cell("gs3d")
    flat = engine.pack_params(cell._packed_tensors(), torch.float32)
    assert flat.numel() == 2 + 125 + 2 * (3 * (2 * 2 + 2) + 2 + 1)
    assert flat[0] == cell.CA.detach() and flat[2 + 62] == cell.W_laplace.weight.detach()[0, 0, 2, 2, 2]


def test_data_loss_selection_matches_oracle_and_python_slicing():
    """Host logic of FusedRCNN.forward_data_loss: list index -> state index through effective_step (GS3D:191-212,
    394-403), against the oracle's restatement and against literally slicing a Python list."""
    from oracle import percnn_oracle as po
    from percnn_b200.cells import data_loss_selection
    cases = [(7, [0, 2, 3, 6], 2, None), (6, range(6), 5, None), (31, range(31), 15, None), (30, range(30), 15, None),
             (9, range(9), 4, 2), (12, [1, 5, 11], 1, None), (5, range(5), 7, None)]
    for step, eff, ts, ff in cases:
        frame_state, sel = data_loss_selection(step, eff, ts, ff)
        outputs = ["h0"] + [f"h{s + 1}" for s in range(step) if s in set(eff)]      # what RCNN.forward returns
        picked = outputs[0:-1:ts]
        if ff is not None:
            picked = picked[:ff]
        want = [int(name[1:]) for name in picked]
        assert [s for s in range(step + 1) if sel[s]] == want
        assert want == po.data_loss_frames(step, eff, ts, ff)
        assert [f"h{i}" for i in frame_state] == outputs
        assert len(sel) == step + 1 and not sel[step] or step in want
    import pytest
    with pytest.raises(ValueError):
        data_loss_selection(3, [], 1)           # outputs = [h0] only: `[0:-1]` is empty
    with pytest.raises(ValueError):
        data_loss_selection(3, range(3), 0)


def test_physics_spec_polynomials_match_the_scripts_formulas():
    """losses.lambda_omega_spec / gray_scott_spec pack the reaction terms of FWD:337-340 / GS2D:321-328 as cubics;
    evaluate the packing on random points against the formulas written out."""
    import numpy as np
    from percnn_b200 import losses
    g = np.random.default_rng(0)
    u, v = g.uniform(-1, 1, 50), g.uniform(-1, 1, 50)
    mono = [np.ones_like(u), u, v, u * u, u * v, v * v, u ** 3, u * u * v, u * v * v, v ** 3]
    ev = lambda c: sum(ci * m for ci, m in zip(c, mono))
    lo = losses.lambda_omega_spec()
    a = u * u + v * v
    np.testing.assert_allclose(ev(lo.poly_u), (1 - a) * u + a * v, atol=1e-14)
    np.testing.assert_allclose(ev(lo.poly_v), -a * u + (1 - a) * v, atol=1e-14)
    assert lo.diff == (0.1, 0.1) and lo.dt == 0.0125 and lo.dx == 0.2
    gs = losses.gray_scott_spec(2e-5, 5e-6, 1 / 25, 3 / 50, 0.5, 0.01)
    np.testing.assert_allclose(ev(gs.poly_u), -u * v * v + (1 / 25) * (1 - u), atol=1e-15)
    np.testing.assert_allclose(ev(gs.poly_v), u * v * v - (1 / 25 + 3 / 50) * v, atol=1e-15)
    from percnn_b200.variants import gs2d, gs3d, lambda_omega_fwd
    assert gs3d.loss_generator().spec.diff == (0.2, 0.1) and gs2d.loss_generator().spec.dx == 0.01
    assert lambda_omega_fwd.loss_generator(dt=0.1, dx=0.5).spec.dt == 0.1


def test_fused_losses_fail_loudly_without_cuda():
    """No CPU fallback anywhere on the product path: CPU tensors raise before any library call."""
    import pytest
    import torch
    from percnn_b200 import losses
    from percnn_b200.variants import gs2d, lambda_omega_fwd
    with pytest.raises(RuntimeError, match="CUDA"):
        losses.physics_loss(torch.zeros(4, 2, 8, 8, dtype=torch.float64), losses.lambda_omega_spec())
    with pytest.raises(RuntimeError, match="CUDA"):
        lambda_omega_fwd.loss_gen(torch.zeros(4, 2, 8, 8, dtype=torch.float64), lambda_omega_fwd.loss_generator())
    cell = gs2d.RCNNCell(input_channels=2, hidden_channels=8, input_kernel_size=5)
    with pytest.raises(RuntimeError, match="(?i)cuda"):
        cell.rollout_data_loss(torch.zeros(1, 2, 16, 16), 3, torch.zeros(1, 2, 8, 8), [True, False, False, False], 2)


def test_fused_upscaler_fails_loudly_without_cuda_and_keeps_the_reference_keys():
    """The drop-in upscalers hold the reference's layers as parameter containers (same state_dict keys, same initial
    values under a seed as the stock modules constructed in the same order) and have no CPU path."""
    import pytest
    import torch
    from percnn_b200 import upscaler as up
    from percnn_b200.variants import burgers_stage1, gs2d, gs3d
    assert list(gs2d.upscaler().state_dict().keys()) == [f"convnet.{i}.{w}" for i in (0, 2, 3) for w in ("weight", "bias")]
    assert list(gs3d.upscaler().state_dict().keys()) == [f"convnet.{i}.{w}" for i in (0, 2, 3) for w in ("weight", "bias")]
    assert list(burgers_stage1.upscaler().state_dict().keys()) == [
        "up0.weight", "up0.bias", "out.weight", "out.bias", "convnet.0.weight", "convnet.0.bias", "convnet.2.weight", "convnet.2.bias"]
    torch.manual_seed(7)
    a = gs2d.upscaler()
    torch.manual_seed(7)
    ref0 = torch.nn.ConvTranspose2d(2, 8, kernel_size=5, padding=2, stride=2, output_padding=1, bias=True)
    assert torch.equal(a.convnet[0].weight, ref0.weight)
    assert [tuple(p.shape) for p in a.up_parameters()] == [(2, 8, 5, 5), (8,), (8, 8, 5, 5), (8,), (2, 8, 1, 1), (2,)]
    with pytest.raises(RuntimeError, match="(?i)cuda"):
        a(torch.zeros(1, 2, 5, 5))
    with pytest.raises(RuntimeError, match="(?i)cuda"):
        up.mse(torch.zeros(4), torch.zeros(4))


def test_stage3_burgers_checkpoint_with_stale_keys_loads():
    """The shipped Stage-3 Burgers checkpoint carries C3_*/C4_* coefficients of an older script version;
    the drop-in drops them with a warning and still insists on every key it does have."""
    import warnings
    from percnn_b200.variants import burgers_stage3
    z = np.load(os.path.join(ROOT, "tests", "golden", "ckpt_bur3_stage3.npz"))
    sd = {k: torch.from_numpy(z[k]) for k in z.files}
    assert "crnn_cell.C3_u" in sd
    model = burgers_stage3.RCNN(input_channels=2, hidden_channels=4, output_channels=2, init_state_low=torch.zeros(1, 2, 5, 5, dtype=torch.float64),
                                input_kernel_size=5, input_stride=1, input_padding=2, step=2, effective_step=[0, 1])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        model.load_state_dict(sd)          # strict
    assert any("stale" in str(x.message) for x in w)
    assert float(model.crnn_cell.nu_u) == float(z["crnn_cell.nu_u"])
    assert torch.equal(model.UpconvBlock.up0.weight.detach(), sd["UpconvBlock.up0.weight"])
    del sd["crnn_cell.C1_u"]
    with pytest.raises(RuntimeError):
        model.load_state_dict(sd)


def test_upscaler_slab_geometry_partitions_the_grid():
    """percnn_upscaler_sizes is host arithmetic (no device needed): for every rank count the per-rank activation tapes
    must cover what that rank's output planes read (2 planes of halo, clipped at the global border) and the output
    extents / parameter counts must match the reference's layers (GS3D:41-56: 2 -> 8 -> 8 -> 2, kernels of 125)."""
    import ctypes
    from percnn_b200 import _lib
    L = _lib.lib()

    def sizes(low, z0=0, nz=0, stride2=1, ndim=3, channels=8, layers=2):
        d = _lib.Upscaler()
        d.ndim, d.dtype, d.channels, d.act, d.layers, d.stride2, d.device = ndim, _lib.F32, channels, 0, layers, stride2, 0
        ext = (1,) + tuple(low) if ndim == 2 else tuple(low)
        for i in range(3):
            d.low_extent[i] = ext[i]
        d.out_z0, d.out_nz, d.out_field_stride = z0, nz, 0
        npar, mid, ws = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_size_t()
        out = (ctypes.c_int64 * 3)()
        rc = L.percnn_upscaler_sizes(ctypes.byref(d), ctypes.byref(npar), ctypes.byref(mid), out, ctypes.byref(ws))
        return rc, npar.value, mid.value, tuple(out), ws.value

    rc, npar, mid, out, ws = sizes((24, 24, 24))
    assert rc == 0 and out == (48, 48, 48)
    assert npar == 2 * 8 * 125 + 8 + 8 * 8 * 125 + 8 + 2 * 8 + 2          # GS3D:45-52
    assert mid == 8 * 48 ** 3 and ws > 0
    for world in (2, 3, 4, 6, 8):
        nz = 48 // world
        for r in range(world):
            rc, _, mid_r, _, _ = sizes((24, 24, 24), z0=r * nz, nz=nz)
            assert rc == 0
            lo, hi = max(r * nz - 2, 0), min(r * nz + nz + 2, 48)          # stride-1 second layer: planes z-2 .. z+2
            assert mid_r == 8 * (hi - lo) * 48 * 48, (world, r)
    # GS2D (two stride-2 layers): 25^2 -> 100^2, and the one-layer Stage-1 net: 50^2 -> 100^2 with 16 channels
    assert sizes((25, 25), stride2=2, ndim=2)[3] == (1, 100, 100)
    rc, npar, mid, out, _ = sizes((50, 50), ndim=2, channels=16, layers=1)
    assert rc == 0 and out == (1, 100, 100) and npar == 2 * 16 * 25 + 16 + 2 * 16 + 2 and mid == 16 * 100 * 100
    # rejected: a plane range that does not lie in the grid, slab ranges for 2-D nets, bad channel counts
    assert sizes((24, 24, 24), z0=40, nz=16)[0] != 0
    assert sizes((25, 25), z0=0, nz=10, ndim=2, stride2=2)[0] != 0
    assert sizes((24, 24, 24), channels=12)[0] != 0
