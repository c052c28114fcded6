"""One GPU, ring of one rank: taped slab rollout + fused-loss adjoint (the cfg5 training step's kernels) per time step."""
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
for shape, T in (((64, 512, 512), 30), ((256, 512, 512), 15)):
    slab = halo.SlabRollout(cell, shape, dev, 0, 1, transport="fused")
    h0 = synthetic_state(shape, 0, shape[0], dev, torch.float32)
    sel = tuple((s % 15 == 0) and s < T for s in range(T + 1))
    tgt = torch.rand((sum(sel), 2, shape[0] // 2, shape[1] // 2, shape[2] // 2), device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    best_f = best_b = 1e30
    for rep in range(3):
        slab.set_state(h0)
        torch.cuda.synchronize()
        ev[0].record()
        tape = slab.rollout_tape(T)
        ev[1].record()
        g_h0, grads = slab.backward(tape, None, loss=(tgt, sel, 2, None))
        ev[2].record()
        torch.cuda.synchronize()
        best_f = min(best_f, ev[0].elapsed_time(ev[1]) / T)
        best_b = min(best_b, ev[1].elapsed_time(ev[2]) / T)
    n = shape[0] * shape[1] * shape[2]
    print(f"{shape} T={T}: taped fwd {best_f*1e3:7.1f} us/step ({n*16/best_f/1e6:6.0f} GB/s)  adjoint {best_b*1e3:7.1f} us/step "
          f"({n*24/best_b/1e6:6.0f} GB/s)", flush=True)
    del slab, tape, g_h0, grads, tgt, h0
    engine.clear_plans()
    torch.cuda.empty_cache()
