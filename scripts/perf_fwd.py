"""Forward step kernel timing at 512^3 (CUDA events around back-to-back launches, short bursts)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import load_gs3d_weights, synthetic_state  # noqa: E402
from percnn_b200 import engine  # noqa: E402
from percnn_b200.variants import gs3d  # noqa: E402

dev = torch.device("cuda:0")
cell = gs3d.RCNNCell(2, 2, 5)
cell.load_state_dict(load_gs3d_weights())
cell = cell.to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
shape = (n, n, n)
plan = engine.get_plan(cell._spec(), shape, dev)
plan.params_load(engine.pack_params(cell._packed_tensors(), torch.float32))
a = synthetic_state(shape, 0, n, dev, torch.float32)
b = torch.empty_like(a)
for steps in (20, 200):
    for _ in range(3):
        plan.step_fwd(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record()
        for _ in range(steps // 2):
            plan.step_fwd(a, b)
            plan.step_fwd(b, a)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    print(f"forward step {n}^3, {steps} back-to-back: {best*1e3:8.1f} us  {n**3*16/best/1e6:8.1f} GB/s", flush=True)
