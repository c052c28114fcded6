"""One GPU: the slab-mode fused-halo kernel as a ring of one rank (its own neighbour) vs the periodic kernel on the
same grid -- isolates the cost of the slab kernel itself (mirror stores, flags, march direction) from NVLink."""
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
steps = int(os.environ.get("STEPS", "300"))


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for shape in [(512, 512, 512), (256, 512, 512), (64, 512, 512), (128, 128, 128), (16, 128, 128)]:
    D = shape[0]
    h0 = synthetic_state(shape, 0, D, dev, torch.float32)
    plan = engine.get_plan(cell._spec(), shape, dev)
    plan.params_load(flat)
    out = torch.empty_like(h0)
    ms_p = timed(lambda: plan.rollout_fwd(h0, steps, h_final=out))
    slab = halo.SlabRollout(cell, shape, dev, 0, 1, transport="fused")
    slab.set_state(h0)
    ms_s = timed(lambda: slab.run(steps))
    print(f"{shape}: periodic {ms_p/steps*1e3:8.1f} us/step   slab-self {ms_s/steps*1e3:8.1f} us/step   "
          f"({shape[0]*shape[1]*shape[2]*16/(ms_s/steps)/1e6:6.0f} GB/s)", flush=True)
    del slab, plan, h0, out
    engine.clear_plans()
    torch.cuda.empty_cache()

# the PLAIN kernel on the ghosted slab layout (no exchange at all): separates "layout" from "slab kernel code"
for shape in [(512, 512, 512), (64, 512, 512), (128, 128, 128)]:
    plan = engine.get_plan(cell._spec(), shape, dev, slab_ghost=True)
    plan.params_load(flat)
    a = torch.zeros(plan.buffer_shape, device=dev)
    b = torch.zeros_like(a)
    a[:, 2:-2] = synthetic_state(shape, 0, shape[0], dev, torch.float32)

    def run():
        x, y = a, b
        for _ in range(steps):
            plan.step_fwd_range(x, y, 0, shape[0])
            x, y = y, x

    ms = timed(run)
    print(f"{shape}: plain kernel on ghost layout {ms/steps*1e3:8.1f} us/step", flush=True)
    del plan, a, b
    engine.clear_plans()
    torch.cuda.empty_cache()
