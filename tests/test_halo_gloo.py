"""world_size-2/3 CPU (gloo) tests of the slab decomposition host logic: ghost-plane exchange pairing (the
world == 2 case where both ring neighbours are the same peer), slab bounds, and that "exchange, then a
radius-2 stencil step on the ghosted slab" reproduces the global periodic step.  The stencil step used
here is the ORACLE (test infrastructure); the product's step kernels are CUDA-only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import percnn_oracle as po
from percnn_b200.halo import exchange_ghosts, slab_bounds
from tests.helpers import load_weights


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, nsteps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        D, H, W = shape
        z0, nz = slab_bounds(D, rank, world)
        params = load_weights("gs3d")
        full = po.ic_gs_3d(shape, seed=2)
        mine = po.ic_gs_3d((nz, H, W), seed=2, z0=z0, z_total=D)         # slab-wise generation == global field
        assert torch.equal(mine, full[:, :, z0:z0 + nz])
        buf = torch.zeros(2, nz + 4, H, W)
        buf[:, 2:nz + 2] = mine[0]
        for _ in range(nsteps):
            for w in exchange_ghosts(buf, nz, rank, world):
                w.wait()
            # ghosts must be the neighbours' boundary planes
            stepped = po.cell_step_torch(buf[None], params, "gs3d")[0]  # periodic in z over the ghosted slab:
            buf[:, 2:nz + 2] = stepped[:, 2:nz + 2]                     # only planes whose stencil stays inside are kept
        ref = full
        for _ in range(nsteps):
            ref = po.cell_step_torch(ref, params, "gs3d")
        err = float((buf[:, 2:nz + 2] - ref[0, :, z0:z0 + nz]).abs().max())
        # ghost check after a final exchange
        for w in exchange_ghosts(buf, nz, rank, world):
            w.wait()
        lo = ref[0, :, (z0 - 2) % D:(z0 - 2) % D + 2]
        hi = ref[0, :, (z0 + nz) % D:(z0 + nz) % D + 2]
        gerr = max(float((buf[:, 0:2] - lo).abs().max()), float((buf[:, nz + 2:] - hi).abs().max()))
        out[rank] = (err, gerr)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_exchange_reproduces_global_step(world):
    shape = (12, 6, 8)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), shape, 3, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        err, gerr = out[r]
        assert err <= 1e-6 and gerr <= 1e-6, (r, err, gerr)


def test_world1_wrap_is_a_local_copy():
    nz = 6
    buf = torch.arange(2 * (nz + 4) * 2 * 3, dtype=torch.float32).view(2, nz + 4, 2, 3)
    exchange_ghosts(buf, nz, 0, 1)
    assert torch.equal(buf[:, 0:2], buf[:, nz:nz + 2]) and torch.equal(buf[:, nz + 2:], buf[:, 2:4])


def test_slab_bounds():
    assert slab_bounds(512, 3, 8) == (192, 64)
    with pytest.raises(ValueError):
        slab_bounds(10, 0, 4)


def _loss_worker(rank, world, port, shape, nsteps, sel, stride, out):
    """Each rank reduces the fused data loss over its own slab (oracle arithmetic, the product's slab_loss_spec for
    the bookkeeping); one all-reduce must give the global strided-subsample MSE, and the per-slab loss gradients
    must tile the global one."""
    from percnn_b200.halo import slab_loss_spec
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D, H, W = shape
        z0, nz = slab_bounds(D, rank, world)
        g = torch.Generator().manual_seed(4)
        states = torch.rand((nsteps + 1, 2, D, H, W), generator=g, dtype=torch.float64)
        frames = [s for s in range(nsteps + 1) if sel[s]]
        truth = torch.rand((len(frames), 2, -(-D // stride), -(-H // stride), -(-W // stride)), generator=g, dtype=torch.float64)
        spec = slab_loss_spec(shape, z0, nz, nsteps, sel, stride)
        want = po.data_loss_np(states.numpy(), truth.numpy(), frames, stride)
        want_g = po.data_loss_grad_np(states.numpy(), truth.numpy(), frames, stride)
        # slab-local sum of squares over the local lattice, divided by the GLOBAL count
        loc = states[:, :, z0:z0 + nz].numpy()
        loc_truth = truth[:, :, z0 // stride:(z0 + nz) // stride].numpy()
        sub = loc[frames][:, :, ::stride, ::stride, ::stride]
        part = torch.tensor([float(((sub - loc_truth) ** 2).sum() / spec.n_total)], dtype=torch.float64)
        dist.all_reduce(part)
        grad_loc = np.zeros_like(loc)
        for i, f in enumerate(frames):
            grad_loc[f][:, ::stride, ::stride, ::stride] = 2.0 / spec.n_total * (sub[i] - loc_truth[i])
        out[rank] = (abs(float(part) - want) / want, float(np.abs(grad_loc - want_g[:, :, z0:z0 + nz]).max()), spec.n_total)
    finally:
        dist.destroy_process_group()


def test_slab_data_loss_partial_sums_add_up_to_the_global_loss():
    shape, nsteps, stride, world = (12, 6, 10), 6, 2, 2
    sel = [(s % 3 == 0) and s < nsteps for s in range(nsteps + 1)]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_loss_worker, args=(world, _free_port(), shape, nsteps, sel, stride, out), nprocs=world, join=True)
    for r in range(world):
        rel, gerr, n_total = out[r]
        assert rel <= 1e-13 and gerr <= 1e-15
        assert n_total == 2 * 2 * 6 * 3 * 5


def test_slab_loss_spec_rejects_misaligned_slabs_and_last_state():
    from percnn_b200.halo import slab_loss_spec
    with pytest.raises(ValueError):
        slab_loss_spec((12, 8, 8), 3, 3, 4, [True, False, False, False, False], 2)       # slab origin off the lattice
    with pytest.raises(ValueError):
        slab_loss_spec((12, 8, 8), 0, 6, 4, [True, False, False], 2)                      # mask length
    with pytest.raises(NotImplementedError):
        slab_loss_spec((12, 8, 8), 0, 6, 4, [True, False, False, False, True], 2)         # h_nsteps selected
    spec = slab_loss_spec((12, 9, 7), 6, 6, 4, [True, False, True, False, False], 3)
    assert spec.n_total == 2 * 2 * 4 * 3 * 3 and spec.stride == 3 and spec.nsel == 2
