"""Few steps of the periodic kernel and of the slab kernel (ring of one) on the same grid, for ncu."""
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
flat = engine.pack_params(cell._packed_tensors(), torch.float32)
shape = tuple(int(x) for x in os.environ.get("SHAPE", "64,512,512").split(","))
h0 = synthetic_state(shape, 0, shape[0], dev, torch.float32)
plan = engine.get_plan(cell._spec(), shape, dev)
plan.params_load(flat)
out = torch.empty_like(h0)
plan.rollout_fwd(h0, 6, h_final=out)
torch.cuda.synchronize()
slab = halo.SlabRollout(cell, shape, dev, 0, 1, transport="fused")
slab.set_state(h0)
slab.run(6)
torch.cuda.synchronize()
