"""2-D cells on shared-memory tiles with temporal blocking (csrc/kernels_tile2d.cuh) against the gather kernels:
same per-cell arithmetic (point_ops.cuh), so every state must agree BIT FOR BIT -- tape mode, emitted frames and
final-state-only mode, step counts that are not multiples of the steps per pass, grids that are not multiples of
the tile, and every forced number of steps per pass."""
import pytest
import torch

from percnn_b200 import engine
from tests.helpers import load_weights, make_cell

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = [("fwd", "fwd", (128, 128)), ("gs2d", "gs2d", (256, 256)), ("gs2d", "gs2d", (100, 100)), ("bur3", None, (512, 512)),
         ("lo3", None, (48, 64)), ("lo3n", None, (20, 24)), ("fwd", "fwd", (36, 52)), ("gs2d", "gs2d", (24, 16))]


def _cell(tag, alias):
    cell = make_cell(tag)
    if alias:
        cell.load_state_dict(load_weights(alias), strict=True)
    return cell.to(DEV)


def _state(shape, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand((1, 2, *shape), generator=g, dtype=torch.float64) - 0.5) * 1.2).to(dtype).to(DEV)


def _run(cell, h0, nsteps, emit):
    with torch.no_grad():
        states = cell.rollout(h0, nsteps)
        traj, fin = cell.rollout_emit(h0, nsteps, emit, want_final=True)
        _, fin2 = cell.rollout_emit(h0, nsteps, [False] * nsteps, want_final=True)
    return states, traj, fin, fin2


@pytest.mark.parametrize("tag,alias,shape", CASES)
@pytest.mark.parametrize("forced_k", [0, 1, 3])
def test_tiled_rollout_is_bitwise_equal_to_gather_kernels(tag, alias, shape, forced_k, monkeypatch):
    nsteps = 13
    emit = [s in (0, 4, 5, 12) for s in range(nsteps)]
    res = {}
    for mode in ("tile", "gather"):
        monkeypatch.delenv("PERCNN_NO_TILE2D", raising=False)
        monkeypatch.delenv("PERCNN_TILE2D_K", raising=False)
        if mode == "gather":
            monkeypatch.setenv("PERCNN_NO_TILE2D", "1")
        elif forced_k:
            monkeypatch.setenv("PERCNN_TILE2D_K", str(forced_k))
        engine.clear_plans()
        cell = _cell(tag, alias)
        h0 = _state(shape, cell.dtype)
        k = engine.get_plan(cell._spec(), shape, torch.device(DEV)).tile2d_steps_per_pass
        assert (k > 0) == (mode == "tile"), (mode, k)
        if mode == "tile" and forced_k:
            assert k == forced_k
        res[mode] = _run(cell, h0, nsteps, emit)
    engine.clear_plans()
    st, tr, f1, f2 = res["tile"]
    gs, gt, g1, g2 = res["gather"]
    assert torch.equal(st, gs)
    assert torch.equal(tr, gt) and torch.equal(f1, g1) and torch.equal(f2, g2)
    assert torch.equal(tr[0], st[1]) and torch.equal(tr[3], st[13]) and torch.equal(f1, st[13]) and torch.equal(f2, st[13])
    assert torch.isfinite(st).all()


@pytest.mark.parametrize("tag,shape", [("bur3", (64, 64)), ("lo3", (40, 48))])
def test_tiled_fp32_physics_cells(tag, shape, monkeypatch):
    """The Stage-3 cells in fp32 (north_star's fp32 target) through the tiled path vs the gather path."""
    out = []
    for no_tile in ("", "1"):
        if no_tile:
            monkeypatch.setenv("PERCNN_NO_TILE2D", "1")
        else:
            monkeypatch.delenv("PERCNN_NO_TILE2D", raising=False)
        engine.clear_plans()
        cell = make_cell(tag).float()
        cell.dtype = torch.float32
        cell = cell.to(DEV)
        with torch.no_grad():
            out.append(cell.rollout(_state(shape, torch.float32, 3), 9))
    engine.clear_plans()
    assert torch.equal(out[0], out[1])
