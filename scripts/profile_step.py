"""Small driver for ncu: a few fused steps of the 3-D Gray-Scott cell at a given size (no timing here)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import load_gs3d_weights, synthetic_state  # noqa: E402
from percnn_b200 import engine  # noqa: E402
from percnn_b200.variants import gs3d  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--bwd", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
cell = gs3d.RCNNCell(2, 2, 5)
cell.load_state_dict(load_gs3d_weights(), strict=True)
cell = cell.to(dev)
shape = (a.n, a.n, a.n)
plan = engine.get_plan(cell._spec(), shape, dev)
plan.params_load(engine.pack_params(cell._packed_tensors(), torch.float32))
h = synthetic_state(shape, 0, a.n, dev, torch.float32)
out = torch.empty_like(h)
plan.rollout_fwd(h, a.steps, h_final=out)
if a.bwd:
    g = torch.ones_like(h)
    gi = torch.empty_like(h)
    plan.param_grads_begin()
    for _ in range(a.steps):
        plan.step_bwd(h, g, gi)
torch.cuda.synchronize()
print("done", float(out.mean()))
