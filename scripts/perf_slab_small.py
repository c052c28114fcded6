"""One GPU, ring of one rank: the persistent small-slab rollout at cfg4-class slab sizes for several K
(PERCNN_SLAB_TB_K: time steps per halo exchange; 1 = per-step hand-shake)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import load_gs3d_weights, synthetic_state  # noqa: E402
from percnn_b200 import engine, halo  # noqa: E402
from percnn_b200.variants import gs3d  # noqa: E402

dev = torch.device("cuda:0")
cell = gs3d.RCNNCell(2, 2, 5)
cell.load_state_dict(load_gs3d_weights())
cell = cell.to(dev)
steps = 500


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for shape in [(16, 128, 128), (32, 128, 128), (64, 128, 128), (128, 128, 128)]:
    h0 = synthetic_state(shape, 0, shape[0], dev, torch.float32)
    line = f"{shape}:"
    for k in (1, 2, 3, 4, 6, 8):
        os.environ["PERCNN_SLAB_TB_K"] = str(k)
        slab = halo.SlabRollout(cell, shape, dev, 0, 1, transport="fused")
        slab.set_state(h0)
        ms = timed(lambda: slab.run(steps))
        line += f"  K={k}: {ms/steps*1e3:6.2f} us/step"
        del slab
        engine.clear_plans()
    print(line, flush=True)
