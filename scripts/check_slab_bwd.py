"""torchrun entry: slab-decomposed training step (taped forward + fused-halo adjoint + gradient all-reduce)
vs single-GPU autograd through the same rollout and loss."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import load_gs3d_weights, synthetic_state  # noqa: E402
from percnn_b200 import engine, halo  # noqa: E402
from percnn_b200.variants import gs3d  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", type=int, nargs=3, default=[64, 48, 128])
ap.add_argument("--steps", type=int, default=6)
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
shape = tuple(a.shape)
D, H, W = shape
cell = gs3d.RCNNCell(2, 2, 5)
cell.load_state_dict(load_gs3d_weights())
cell = cell.to(dev)
slab = halo.SlabRollout(cell, shape, dev, rank, world, transport="fused")
full = synthetic_state(shape, 0, D, dev, torch.float32, seed=3)
g = torch.Generator(device=dev).manual_seed(7)
target = torch.rand((a.steps + 1, 2, D, H, W), generator=g, device=dev)        # same on every rank (same seed)
z0, nz = slab.z0, slab.nz

slab.set_state(full[:, z0:z0 + nz])
tape = slab.rollout_tape(a.steps)
s = tape.detach().clone().requires_grad_(True)
loss_local = ((s[:, :, 2:nz + 2, ::2, ::2] - target[:, :, z0:z0 + nz, ::2, ::2]) ** 2).sum() / target[:, :, :, ::2, ::2].numel()
loss_local.backward()
g_h0, grads = slab.backward(tape, s.grad)
loss = loss_local.detach().clone()
dist.all_reduce(loss)

# reference: single GPU autograd through the fused rollout on the whole grid
h0 = full[None].clone().requires_grad_(True)
states = cell.rollout(h0, a.steps)
ref_loss = ((states[:, :, :, ::2, ::2] - target[:, :, :, ::2, ::2]) ** 2).mean()
ref_loss.backward()
ref_flat = engine.pack_params([p.grad if p.grad is not None else torch.zeros_like(p) for p in cell._packed_tensors()], torch.float32)


def rel(a_, b_):
    return float((a_.double() - b_.double()).norm() / b_.double().norm().clamp_min(1e-300))


e_h0 = rel(g_h0, h0.grad[0, :, z0:z0 + nz])
e_p = rel(grads, ref_flat)
e_loss = abs(float(loss) - float(ref_loss)) / abs(float(ref_loss))
# ---- the same training step with the FUSED data loss (no dense gradient tape): frames 0, 2, 4, ... < steps, stride 2
sel = [(t % 2 == 0) and t < a.steps for t in range(a.steps + 1)]
frames = [t for t in range(a.steps + 1) if sel[t]]
truth_sub = target[frames][:, :, ::2, ::2, ::2].contiguous()                    # global low-res truth
slab.set_state(full[:, z0:z0 + nz])
tape = slab.rollout_tape(a.steps)
local_truth = truth_sub[:, :, z0 // 2:(z0 + nz) // 2].contiguous()
f_loss = slab.data_loss(tape, local_truth, sel, 2)
f_g_h0, f_grads = slab.backward(tape, None, loss=(local_truth, sel, 2, torch.tensor(3.0, device=dev)))
for p_ in cell.parameters():
    p_.grad = None
h0f = full[None].clone().requires_grad_(True)
_, ref_f_loss = cell.rollout_data_loss(h0f, a.steps, truth_sub, sel, 2)
(3.0 * ref_f_loss).backward()
ref_f_flat = engine.pack_params([p.grad if p.grad is not None else torch.zeros_like(p) for p in cell._packed_tensors()], torch.float32)
e_fused = max(rel(f_g_h0, h0f.grad[0, :, z0:z0 + nz]), rel(f_grads, ref_f_flat),
              abs(float(f_loss) - float(ref_f_loss)) / abs(float(ref_f_loss)))

worst = torch.tensor([e_h0, e_p, max(e_loss, e_fused), float(slab.error_word())], device=dev, dtype=torch.float64)
dist.all_reduce(worst, op=dist.ReduceOp.MAX)
if rank == 0:
    ok = worst[0] < 1e-5 and worst[1] < 1e-5 and worst[2] < 1e-5 and worst[3] == 0
    print(f"SLAB_BWD_CHECK world={world} shape={shape} steps={a.steps}: loss/fused-loss rel {worst[2]:.2e}  dL/dh0 rel {worst[0]:.2e}  "
          f"param-grad rel {worst[1]:.2e}  device_error_word={int(worst[3])}  ok={bool(ok)}", flush=True)
dist.barrier()
dist.destroy_process_group()
