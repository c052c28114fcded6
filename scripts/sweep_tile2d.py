"""Sweep steps-per-pass and tile shapes of the 2-D tiled kernel (PERCNN_TILE2D_K / _TH / _TW) on the BASELINE 2-D configs."""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from percnn_b200 import engine  # noqa: E402
from tests.helpers import load_weights, make_cell  # noqa: E402

DEV = "cuda:0"


def run(tag, alias, n, steps):
    engine.clear_plans()
    cell = make_cell(tag)
    if alias:
        cell.load_state_dict(load_weights(alias))
    cell = cell.to(DEV)
    g = torch.Generator().manual_seed(0)
    h0 = ((torch.rand((1, 2, n, n), generator=g, dtype=torch.float64) - 0.5)).to(cell.dtype).to(DEV)
    k = engine.get_plan(cell._spec(), (n, n), torch.device(DEV)).tile2d_steps_per_pass
    emit = [False] * steps
    with torch.no_grad():
        cell.rollout_emit(h0, steps, emit, want_final=True)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            cell.rollout_emit(h0, steps, emit, want_final=True)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    return k, best / steps * 1e3


cases = {"cfg1": ("fwd", "fwd", 128, 200, [(8, 16), (16, 16), (16, 32), (10, 32), (32, 32), (13, 16)]),
         "cfg2": ("gs2d", "gs2d", 256, 1000, [(16, 32), (32, 16), (22, 32), (32, 32), (16, 64), (32, 64), (26, 32)]),
         "cfg3ii": ("bur3", None, 512, 40, [(32, 64), (64, 32), (43, 64), (64, 64), (32, 128), (52, 64)])}
for name, (tag, alias, n, steps, tiles) in cases.items():
    for v in ("PERCNN_TILE2D_K", "PERCNN_TILE2D_TH", "PERCNN_TILE2D_TW"):
        os.environ.pop(v, None)
    os.environ["PERCNN_TILE2D_VERBOSE"] = "1"
    k, us = run(tag, alias, n, steps)
    os.environ.pop("PERCNN_TILE2D_VERBOSE")
    print(f"{name} model choice K={k}: {us:.2f} us/step", flush=True)
    for K, (th, tw) in itertools.product([1, 2, 3, 4, 6, 8], tiles):
        os.environ.update(PERCNN_TILE2D_K=str(K), PERCNN_TILE2D_TH=str(th), PERCNN_TILE2D_TW=str(tw))
        k, us = run(tag, alias, n, steps)
        if k:
            print(f"{name} K={K} tile {th}x{tw}: {us:.2f} us/step", flush=True)
