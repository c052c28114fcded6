"""ncu driver: a few forward and adjoint steps of the 5x5 Pi-block cell at 512^2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from percnn_b200 import engine  # noqa: E402
from tests.helpers import load_weights, make_cell  # noqa: E402

dev = torch.device("cuda:0")
cell = make_cell("bur1")
cell.load_state_dict(load_weights("bur1"))
cell = cell.to(dev)
shape = (512, 512)
plan = engine.get_plan(cell._spec(), shape, dev)
plan.params_load(engine.pack_params(cell._packed_tensors(), torch.float32))
g = torch.Generator().manual_seed(0)
h = (torch.rand((2, *shape), generator=g) - 0.5).to(dev)
out = torch.empty_like(h)
gi = torch.empty_like(h)
go = torch.randn_like(h)
plan.param_grads_begin()
for _ in range(4):
    plan.step_fwd(h, out)
    plan.step_bwd(h, go, gi)
torch.cuda.synchronize()
print("ok")
