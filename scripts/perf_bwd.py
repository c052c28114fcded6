"""Adjoint step kernel timing at a given size (CUDA events around back-to-back launches)."""
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
for n in (128, 256, 512):
    shape = (n, n, n)
    plan = engine.get_plan(cell._spec(), shape, dev)
    plan.params_load(engine.pack_params(cell._packed_tensors(), torch.float32))
    h = synthetic_state(shape, 0, n, dev, torch.float32)
    g = torch.randn_like(h)
    ga = torch.randn_like(h)
    gi = torch.empty_like(h)
    plan.param_grads_begin()
    for with_add in (False, True):
        steps = 20 if n == 512 else 100
        for _ in range(3):
            plan.step_bwd(h, g, gi, ga if with_add else None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for rep in range(3):
            e0.record()
            for _ in range(steps):
                plan.step_bwd(h, g, gi, ga if with_add else None)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
        byts = n ** 3 * (32 if with_add else 24)
        print(f"adjoint step {n}^3 g_add={with_add}: {best*1e3:8.1f} us  {byts/best/1e6:8.1f} GB/s algorithmic ({32 if with_add else 24} B/cell)", flush=True)
    del h, g, ga, gi
