"""torchrun entry: slab-decomposed rollout vs the single-GPU rollout of the same global field (bitwise).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/check_slab.py --shape 64 64 128 --steps 7 --transport nccl
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import load_gs3d_weights, synthetic_state  # noqa: E402
from percnn_b200 import engine, halo  # noqa: E402
from percnn_b200.variants import gs3d  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", type=int, nargs=3, default=[64, 64, 128])
ap.add_argument("--steps", type=int, default=7)
ap.add_argument("--transport", default="nccl")
ap.add_argument("--time-steps", type=int, default=0)
ap.add_argument("--repeat", type=int, default=1)
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
shape = tuple(a.shape)
cell = gs3d.RCNNCell(2, 2, 5)
cell.load_state_dict(load_gs3d_weights())
cell = cell.to(dev)
slab = halo.SlabRollout(cell, shape, dev, rank, world, transport=a.transport)
full = synthetic_state(shape, 0, shape[0], dev, torch.float32, seed=3)       # every rank builds the same global field
from percnn_b200 import _lib  # noqa: E402
with torch.no_grad():
    if a.transport == "fused" and slab.plan.slab_persistent and a.steps >= 2:
        cell._flags = _lib.FLAG_NO_TMA       # small slabs: persistent kernel = the gather kernel's arithmetic
    ref = cell.rollout(full[None], a.steps)[-1]
    cell._flags = 0
ok, err = True, 0.0
for rep in range(a.repeat):
    slab.set_state(full[:, slab.z0:slab.z0 + slab.nz])
    slab.run(a.steps)
    torch.cuda.synchronize()
    mine = slab.interior()
    ok = ok and torch.equal(mine, ref[:, slab.z0:slab.z0 + slab.nz]) and slab.error_word() == 0
    err = max(err, float((mine - ref[:, slab.z0:slab.z0 + slab.nz]).abs().max()))
    bad = (mine != ref[:, slab.z0:slab.z0 + slab.nz])
    if bool(bad.any()):
        zs = bad.any(dim=0).flatten(1).any(dim=1).nonzero().flatten().tolist()
        ys = bad.any(dim=0).any(dim=0).any(dim=1).nonzero().flatten().tolist()
        xs = bad.any(dim=0).any(dim=0).any(dim=0).nonzero().flatten().tolist()
        fs = bad.flatten(1).any(dim=1).nonzero().flatten().tolist()
        print(f"MISMATCH rank={rank} rep={rep} cells={int(bad.sum())} fields={fs} local_planes={zs[:12]}{'...' if len(zs) > 12 else ''} "
              f"rows={ys[:10]}..{ys[-2:]} cols={xs[:6]}..{xs[-3:]} ncols={len(xs)} err_word={slab.error_word()}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"SLAB_CHECK transport={a.transport} world={world} shape={shape} steps={a.steps} repeats={a.repeat} bitwise_equal={bool(flag.item())} max_abs_err_rank0={err:.3e}", flush=True)
if a.time_steps:
    slab.run(20)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    slab.run(a.time_steps)
    e1.record()
    torch.cuda.synchronize()
    cpu_ms = (time.perf_counter() - t0) * 1e3
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        n = shape[0] * shape[1] * shape[2]
        print(f"SLAB_TIME transport={a.transport} world={world} shape={shape}: {t.item()/a.time_steps*1e3:.1f} us/step "
              f"(cpu enqueue+wait {cpu_ms/a.time_steps*1e3:.1f} us/step) -> {n*16/(t.item()/a.time_steps)/1e6:.0f} GB/s aggregate", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
