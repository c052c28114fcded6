"""Where does a cfg1 training epoch (lambda-omega 128^2 fp64, 200 steps, physics loss) spend its time?"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from percnn_b200 import engine  # noqa: E402
from percnn_b200.variants import lambda_omega_fwd as mod  # noqa: E402
from tests.helpers import load_weights, make_cell  # noqa: E402

dev = torch.device("cuda:0")
steps = 200
cell = make_cell("fwd")
cell.load_state_dict(load_weights("fwd"))
cell = cell.to(dev)
h0 = (torch.rand((1, 2, 128, 128), dtype=torch.float64, generator=torch.Generator().manual_seed(0)) * 1.6 - 0.8).to(dev)
gen = mod.loss_generator()


def timed(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    best_gpu, best_wall = 1e30, 1e30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        t0 = time.perf_counter(); e0.record(); fn(); e1.record(); t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        best_gpu = min(best_gpu, e0.elapsed_time(e1)); best_wall = min(best_wall, t_issue * 1e3)
    print(f"{name:48s} gpu {best_gpu:8.3f} ms   host issue {best_wall:8.3f} ms", flush=True)


with torch.no_grad():
    timed("rollout no-grad (final only)", lambda: cell.rollout_emit(h0, steps, [False] * steps, want_final=True))
timed("rollout with tape (autograd fwd)", lambda: cell.rollout(h0, steps))
out = cell.rollout(h0, steps).detach()
timed("physics loss fwd only", lambda: mod.loss_gen(out, gen))
outg = out.clone().requires_grad_(True)
timed("physics loss fwd+bwd", lambda: mod.loss_gen(outg, gen).backward())


def full():
    for p in cell.parameters():
        p.grad = None
    o = cell.rollout(h0, steps)
    mod.loss_gen(o, gen).backward()


timed("epoch: rollout + loss + backward", full)
plan = cell._plan(h0)
flat = engine.pack_params(cell._packed_tensors(), torch.float64)
plan.params_load(flat)
tape = cell.rollout(h0, steps).detach()
g = torch.randn_like(tape)
timed("rollout_bwd alone (dense g)", lambda: plan.rollout_bwd(flat, tape, g, [True] * (steps + 1), steps))
timed("rollout_bwd alone (no g_add)", lambda: plan.rollout_bwd(flat, tape, g[-1:].contiguous(), [False] * steps + [True], steps))
gi = torch.empty_like(tape[0])
plan.param_grads_begin()
timed("200 x step_bwd via ctypes", lambda: [plan.step_bwd(tape[3], g[0], gi) for _ in range(steps)])
