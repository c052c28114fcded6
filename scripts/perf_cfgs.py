"""Per-config timings (BASELINE.json configs 1-4) on one GPU: CUDA events around whole rollouts."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests.helpers import load_weights, make_cell  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def state(shape, dtype, lo=0.1, hi=0.9, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand((1, 2, *shape), generator=g, dtype=torch.float64) * (hi - lo) + lo).to(dtype).to(dev)


def fwd_case(name, tag, alias, shape, nsteps, lo=0.1, hi=0.9):
    cell = make_cell(tag)
    if alias:
        cell.load_state_dict(load_weights(alias))
    cell = cell.to(dev)
    h0 = state(shape, cell.dtype, lo, hi)
    emit = [False] * nsteps
    with torch.no_grad():
        ms = timed(lambda: cell.rollout_emit(h0, nsteps, emit, want_final=True))
    ncell = int(np.prod(shape))
    esz = 4 if cell.dtype == torch.float32 else 8
    print(f"{name:34s} fwd  {nsteps:5d} steps: {ms:9.3f} ms  {ms/nsteps*1e3:8.2f} us/step  {nsteps/ms*1e3:10.0f} steps/s  "
          f"{ncell*nsteps/ms/1e6:9.2f} Gcell-steps/s  eff {ncell*4*esz*nsteps/ms/1e6:8.1f} GB/s", flush=True)


def train_case(name, tag, alias, shape, nsteps, lo=0.1, hi=0.9):
    cell = make_cell(tag)
    if alias:
        cell.load_state_dict(load_weights(alias))
    cell = cell.to(dev)
    h0 = state(shape, cell.dtype, lo, hi).requires_grad_(True)
    target = state(shape, cell.dtype, lo, hi, seed=2)

    def step():
        for p in cell.parameters():
            p.grad = None
        h0.grad = None
        out = cell.rollout(h0, nsteps)
        sl = (slice(None), slice(None)) + (slice(None, None, 2),) * len(shape)
        loss = ((out[0:-1:5][sl] - target[sl]) ** 2).mean()
        loss.backward()
    try:
        ms = timed(step, reps=3)
        ncell = int(np.prod(shape))
        print(f"{name:34s} f+b  {nsteps:5d} steps: {ms:9.3f} ms  {ms/nsteps*1e3:8.2f} us/step  {nsteps/ms*1e3:10.0f} steps/s  "
              f"{ncell*nsteps/ms/1e6:9.2f} Gcell-steps/s", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{name:34s} f+b  FAILED: {str(e)[:120]}", flush=True)


def train_fused_loss_case(name, tag, alias, shape, nsteps, lo=0.1, hi=0.9, tstride=5):
    """Same training step as train_case but with the fused data loss (no dense gradient tape)."""
    cell = make_cell(tag)
    if alias:
        cell.load_state_dict(load_weights(alias))
    cell = cell.to(dev)
    h0 = state(shape, cell.dtype, lo, hi).requires_grad_(True)
    sel = [(s % tstride == 0) and s < nsteps for s in range(nsteps + 1)]
    low = tuple((n + 1) // 2 for n in shape)
    target = state(low, cell.dtype, lo, hi, seed=2).expand(sum(sel), *([-1] * (1 + len(shape)))).contiguous()

    def step():
        for p in cell.parameters():
            p.grad = None
        h0.grad = None
        _, loss = cell.rollout_data_loss(h0, nsteps, target, sel, 2)
        loss.backward()
    ms = timed(step, reps=3)
    ncell = int(np.prod(shape))
    print(f"{name:34s} f+b* {nsteps:5d} steps: {ms:9.3f} ms  {ms/nsteps*1e3:8.2f} us/step  {nsteps/ms*1e3:10.0f} steps/s  "
          f"{ncell*nsteps/ms/1e6:9.2f} Gcell-steps/s  (fused data loss)", flush=True)


def train_phys_case(name, tag, alias, shape, nsteps, lo, hi):
    """FWD:366-373: one epoch = rollout + physics-residual loss over the whole trajectory + backward."""
    from percnn_b200.variants import gs2d, lambda_omega_fwd
    mod = {"fwd": lambda_omega_fwd, "gs2d": gs2d}[tag]
    cell = make_cell(tag)
    if alias:
        cell.load_state_dict(load_weights(alias))
    cell = cell.to(dev)
    h0 = state(shape, cell.dtype, lo, hi)
    gen = mod.loss_generator()

    def step():
        for p in cell.parameters():
            p.grad = None
        out = cell.rollout(h0, nsteps)
        mod.loss_gen(out, gen).backward()
    ms = timed(step, reps=3)
    print(f"{name:34s} f+b  {nsteps:5d} steps: {ms:9.3f} ms  {ms/nsteps*1e3:8.2f} us/step  (rollout + fused physics loss + adjoint)", flush=True)


ONLY = os.environ.get("PERF_ONLY", "")     # e.g. PERF_ONLY=cfg3i: run only the cases whose name contains the string


def _run(fn, name, *a):
    if ONLY in name:
        fn(name, *a)


_run(fwd_case, "cfg1 lambda-omega 128^2 fp64", "fwd", "fwd", (128, 128), 200, -0.8, 0.8)
_run(train_phys_case, "cfg1 lambda-omega 128^2 fp64 epoch", "fwd", "fwd", (128, 128), 200, -0.8, 0.8)
_run(fwd_case, "cfg2 GS 256^2 fp32", "gs2d", "gs2d", (256, 256), 1000)
_run(fwd_case, "cfg3i Burgers k5 512^2 fp32", "bur1", "bur1", (512, 512), 40, -0.5, 0.5)
_run(train_case, "cfg3i Burgers k5 512^2 fp32", "bur1", "bur1", (512, 512), 40, -0.5, 0.5)
_run(fwd_case, "cfg3ii Burgers phys 512^2 fp64", "bur3", None, (512, 512), 40, -0.5, 0.5)
_run(train_case, "cfg3ii Burgers phys 512^2 fp64", "bur3", None, (512, 512), 40, -0.5, 0.5)
_run(fwd_case, "cfg4 GS3D 128^3 fp32", "gs3d", "gs3d", (128, 128, 128), 500)
_run(train_case, "GS3D 128^3 fp32 (TMA adjoint)", "gs3d", "gs3d", (128, 128, 128), 20)
_run(train_case, "GS3D 256^3 fp32 (TMA adjoint)", "gs3d", "gs3d", (256, 256, 256), 20)
_run(train_fused_loss_case, "GS3D 256^3 fp32 (TMA adjoint)", "gs3d", "gs3d", (256, 256, 256), 20)
_run(train_case, "GS2D 256^2 fp32", "gs2d", "gs2d", (256, 256), 200)
_run(fwd_case, "ref-size GS3D 48^3", "gs3d", "gs3d", (48, 48, 48), 300)
_run(fwd_case, "ref-size GS2D 100^2", "gs2d", "gs2d", (100, 100), 400)
