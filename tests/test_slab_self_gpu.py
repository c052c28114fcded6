"""The slab-mode (multi-GPU) kernels on ONE GPU: a ring of one rank is its own neighbour, so the fused-halo kernels
mirror their boundary planes into their own ghost planes and wait on their own flags.  This runs exactly the
kernels of the multi-GPU path (single z-march, alternating direction, in-kernel flags, C-side rollout loops) on
the one-GPU test box; results must equal the periodic single-GPU kernels bit for bit."""
import pytest
import torch

from percnn_b200 import _lib, engine, halo
from tests.helpers import load_weights, make_cell

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _cell():
    cell = make_cell("gs3d")
    cell.load_state_dict(load_weights("gs3d"), strict=True)
    return cell.to(DEV)


def _state(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand((2, *shape), generator=g) * 0.8 + 0.1).to(DEV)


@pytest.mark.parametrize("shape,steps", [((40, 48, 256), 9), ((4, 16, 128), 6), ((7, 37, 128), 5), ((128, 128, 128), 8),
                                         ((64, 512, 512), 3)])
def test_self_ring_forward_is_bitwise_equal_to_periodic_kernel(shape, steps):
    cell = _cell()
    h0 = _state(shape, 3)
    slab = halo.SlabRollout(cell, shape, DEV, 0, 1, transport="fused")
    assert slab.transport == "fused"
    with torch.no_grad():
        tma_ref = cell.rollout(h0[None], steps)[-1]
        ref = tma_ref
        if slab.plan.slab_persistent:
            # small slabs run the whole rollout as one persistent kernel built on the gather kernel: bit-identical
            # to the single-GPU gather kernel, and within rounding of the TMA z-march
            cell._flags = _lib.FLAG_NO_TMA
            ref = cell.rollout(h0[None], steps)[-1]
            cell._flags = 0
            assert float((ref - tma_ref).abs().max() / tma_ref.abs().max()) <= 2e-6
    for rep in range(2):                       # the second run starts on the other march direction when steps is odd
        slab.set_state(h0)
        slab.run(steps)
        assert torch.equal(slab.interior(), ref), (rep, float((slab.interior() - ref).abs().max()))
    assert slab.error_word() == 0


@pytest.mark.parametrize("shape", [(24, 48, 128), (64, 64, 256)])
def test_self_ring_training_step_matches_periodic_adjoint(shape):
    cell = _cell()
    h0 = _state(shape, 4)
    T = 5
    sel = (True, False, True, False, True, False)
    flat = engine.pack_params(cell._packed_tensors(), torch.float32)
    plan1 = engine.get_plan(cell._spec(), shape, DEV)
    plan1.params_load(flat)
    spec = engine.DataLossSpec(sel=sel, stride=2)
    g = torch.Generator(device=DEV).manual_seed(9)
    tgt = torch.rand((3, *plan1.lowres_shape(2)), generator=g, device=DEV)
    tape1 = torch.empty((T + 1, *plan1.buffer_shape), device=DEV)
    plan1.rollout_fwd(h0, T, tape=tape1)
    loss1 = plan1.data_loss_fwd(tape1, T, spec, tgt)
    g_h0_ref, g_flat_ref = plan1.rollout_bwd_loss(flat, tape1, T, spec, tgt)

    slab = halo.SlabRollout(cell, shape, DEV, 0, 1, transport="fused")
    slab.set_state(h0)
    tape = slab.rollout_tape(T)
    nz = shape[0]
    assert torch.equal(tape[:, :, 2:nz + 2], tape1)
    loss = slab.data_loss(tape, tgt, sel, 2)
    g_h0, grads = slab.backward(tape, None, loss=(tgt, sel, 2, None))
    assert abs(float(loss) - float(loss1)) <= 1e-6 * abs(float(loss1))
    assert torch.equal(g_h0, g_h0_ref)
    rel = float((grads.double() - g_flat_ref.double()).norm() / g_flat_ref.double().norm())
    assert rel <= 1e-6, rel
    # dense gradient tape instead of the fused loss
    w = torch.randn(tape.shape, generator=torch.Generator(device=DEV).manual_seed(2), device=DEV)
    slab.set_state(h0)
    tape = slab.rollout_tape(T)
    g_h0d, gradsd = slab.backward(tape, w)
    g_ref, gp_ref = plan1.rollout_bwd(flat, tape1, w[:, :, 2:nz + 2].contiguous(), [True] * (T + 1), T)
    assert torch.equal(g_h0d, g_ref)
    assert float((gradsd.double() - gp_ref.double()).norm() / gp_ref.double().norm()) <= 1e-6


def test_self_ring_initial_state_from_the_sharded_upscaler():
    """GS3D:186 on a slab: the rank generates its planes of h0 straight into the slab buffer (periodic ghosts come from
    the exchange, although the transposed convs themselves pad with zeros) and reduces the upscaler's parameter
    gradient from dL/dh0 with exchanged ghost planes."""
    from percnn_b200 import upscaler as up
    from percnn_b200.variants import gs3d
    torch.manual_seed(8)
    cell = _cell()
    ups = gs3d.upscaler().to(DEV)
    low = torch.rand((1, 2, 8, 8, 64), device=DEV) * 0.8 + 0.1
    shape = (16, 16, 128)
    want = ups(low)
    g = torch.rand((2, *shape), device=DEV) - 0.5
    ups.zero_grad()
    (want[0] * g).sum().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in ups.up_parameters()])
    slab = halo.SlabRollout(cell, shape, DEV, 0, 1, transport="fused")
    slab.set_state_from_upscaler(ups, low)
    assert torch.equal(slab.interior(), want[0].detach())
    b = slab.bufs[slab.cur]
    assert torch.equal(b[:, 0:2], b[:, 16:18]) and torch.equal(b[:, 18:20], b[:, 2:4])     # periodic ghosts of the state
    gp = slab.upscaler_backward(g)
    assert float((gp - ref).abs().max()) <= 1e-6 * float(ref.abs().max())


@pytest.mark.parametrize("shape,k,steps", [((16, 16, 128), 4, 11), ((16, 16, 128), 3, 2), ((8, 32, 128), 4, 9), ((9, 16, 128), 2, 6),
                                           ((40, 48, 256), 4, 10), ((16, 128, 128), 8, 18)])
def test_time_blocked_slab_rollout_is_bitwise_equal(shape, k, steps, monkeypatch):
    """Communication-avoiding persistent rollout (2K ghost planes every K steps) against the single-GPU gather kernel,
    for K that divides / does not divide the step count, K = D/2, and back-to-back calls (odd and even lengths) that
    hand the standard buffers, flags and epochs over to each other and to the per-step path."""
    monkeypatch.setenv("PERCNN_SLAB_TB_K", str(k))
    cell = _cell()
    h0 = _state(shape, 6)
    slab = halo.SlabRollout(cell, shape, DEV, 0, 1, transport="fused")
    assert slab.plan.slab_persistent and slab._wide is not None and slab._wide[3] == min(k, shape[0] // 2)
    assert slab.describe()["time_blocking"]["steps_per_exchange"] == min(k, shape[0] // 2)
    with torch.no_grad():
        cell._flags = _lib.FLAG_NO_TMA
        ref = cell.rollout(h0[None], 2 * steps + 1)
        cell._flags = 0
    slab.set_state(h0)
    slab.run(steps)
    assert torch.equal(slab.interior(), ref[steps]), float((slab.interior() - ref[steps]).abs().max())
    slab.run(steps)                  # back to back: the standard buffers, flags and epochs were handed over intact
    assert torch.equal(slab.interior(), ref[2 * steps])
    b = slab.bufs[slab.cur]
    nz = shape[0]
    assert torch.equal(b[:, 0:2], b[:, nz:nz + 2]) and torch.equal(b[:, nz + 2:nz + 4], b[:, 2:4])   # ghosts valid afterwards
    slab.run(1)                      # a single step takes the per-step TMA kernel (same arithmetic up to rounding)
    err = float((slab.interior() - ref[2 * steps + 1]).abs().max() / ref[2 * steps + 1].abs().max())
    assert err <= 2e-6, err
    assert slab.error_word() == 0


def test_host_upload_and_download_of_a_slab():
    """set_state_from_host / interior_to_host (the end-to-end leg of bench.py at N > 1): same state and ghosts as set_state."""
    cell = _cell()
    shape = (24, 32, 128)
    h0 = _state(shape, 12)
    slab = halo.SlabRollout(cell, shape, DEV, 0, 1, transport="fused")
    host = h0.cpu().pin_memory()
    slab.set_state_from_host(host)
    assert torch.equal(slab.interior(), h0)
    b = slab.bufs[slab.cur]
    assert torch.equal(b[:, 0:2], h0[:, -2:]) and torch.equal(b[:, 26:28], h0[:, 0:2])
    slab.run(4)
    out = torch.empty_like(host).pin_memory()
    slab.interior_to_host(out)
    torch.cuda.synchronize()
    assert torch.equal(out, slab.interior().cpu())
    slab2 = halo.SlabRollout(cell, shape, DEV, 0, 1, transport="fused")
    slab2.set_state(h0)
    slab2.run(4)
    assert torch.equal(out, slab2.interior().cpu())
