"""ncu driver: one training epoch of the forward-simulation script's configuration (cfg1: lambda-omega 128^2 fp64,
200 steps, physics loss), for a per-kernel launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from percnn_b200.variants import lambda_omega_fwd as mod  # noqa: E402
from tests.helpers import load_weights, make_cell  # noqa: E402

dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cell = make_cell("fwd")
cell.load_state_dict(load_weights("fwd"))
cell = cell.to(dev)
h0 = (torch.rand((1, 2, 128, 128), dtype=torch.float64, generator=torch.Generator().manual_seed(0)) * 1.6 - 0.8).to(dev)
gen = mod.loss_generator()
for _ in range(2):
    for p in cell.parameters():
        p.grad = None
    out = cell.rollout(h0, steps)
    mod.loss_gen(out, gen).backward()
torch.cuda.synchronize()
print("ok")
