#!/usr/bin/env python
"""Headline benchmark: timesteps/sec of the 3-D Gray-Scott Pi-block rollout (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload gs3d_512|gs3d_128]

A bench "step" is ONE full rollout (ROLLOUT_STEPS = 500 fused time steps, cfg4's rollout length) of the
V-GS3D cell (k=1, hc=2, fp32, shipped-checkpoint weights) on a 512^3 periodic grid (cfg5's grid; 1 GiB of
state, 2 GiB ping-pong working set > L2 so every step streams from HBM).  With N > 1 the grid is
slab-decomposed along D over N ranks (strong scaling, NCCL / peer-memory halo exchange of 2 ghost planes
per side per step, overlapped with the interior kernel).

value   = timesteps/sec with the state resident in HBM (CUDA events, max over ranks)
e2e     = same metric through the host-buffer C-ABI call percnn_rollout_fwd_host (H2D of parameters and
          initial state from pinned memory + D2H of the final state inside the timed region)
roofline= algorithmic bytes (16 B/cell/step fp32) / event time per step kernel vs MEASURED_PEAKS.json
cpu_baseline / --impl reference = the oracle port of the reference's CPU PyTorch op sequence (the
          reference is Python and cannot travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROLLOUT_STEPS = 500
WORKLOADS = {
    "gs3d_512": dict(shape=(512, 512, 512), desc="V-GS3D k=1 hc=2 fp32, 512^3 periodic, 500-step forward rollout (cfg5 grid, cfg4 rollout)"),
    "gs3d_256": dict(shape=(256, 256, 256), desc="V-GS3D k=1 hc=2 fp32, 256^3 periodic, 500-step forward rollout"),
    "gs3d_128": dict(shape=(128, 128, 128), desc="V-GS3D k=1 hc=2 fp32, 128^3 periodic, 500-step forward rollout (cfg4; L2-resident)"),
}
BYTES_PER_CELL_STEP = 16  # fp32, 2 fields, read once + write once


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_gs3d_weights():
    import numpy as np
    import torch
    z = np.load(os.path.join(ROOT, "tests", "golden", "weights_gs3d.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def synthetic_state(shape, z0, nz, device, dtype, seed=0):
    """u=1, v=0, centre cube (half-width N/8) u=.5 v=.25, + 0.01 uniform noise (SURVEY 8d cfg4/5 recipe).
    Generated plane-block-wise on `device` for the global planes [z0, z0+nz)."""
    import torch
    D, H, W = shape
    g = torch.Generator(device=device).manual_seed(seed * 1000003 + z0)
    z = torch.arange(z0, z0 + nz, device=device).view(-1, 1, 1)
    y = torch.arange(H, device=device).view(1, -1, 1)
    x = torch.arange(W, device=device).view(1, 1, -1)
    inside = ((z - D // 2).abs() < D // 8) & ((y - H // 2).abs() < H // 8) & ((x - W // 2).abs() < W // 8)
    u = torch.where(inside, 0.5, 1.0).to(dtype)
    v = torch.where(inside, 0.25, 0.0).to(dtype)
    h = torch.stack((u, v))
    h += 0.01 * (torch.rand(h.shape, generator=g, device=device, dtype=dtype) - 0.5)
    return h


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for n, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU PyTorch path
# ---------------------------------------------------------------------------------------------

def cpu_reference_rate(shape, budget_s=12.0, steps=1, warmup=1):
    """cells*steps/sec of the reference op sequence on the host cores for a bounded slab sample.

    The reference cell pads periodically on every axis, so a [d, H, W] slab of the 512^3 grid is a valid,
    smaller instance of the same computation (same per-cell work, full H x W planes)."""
    import torch
    from oracle import percnn_oracle as po
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = load_gs3d_weights()
    D, H, W = shape
    d = min(D, 4)
    target = budget_s / max(1, steps + warmup + 2)   # seconds one sampled step may take
    with torch.no_grad():
        while True:  # grow the slab until one step costs about `target` (the CPU rate is not linear in depth)
            h = po.ic_gs_3d((d, H, W), seed=0, z_total=D)
            t0 = time.perf_counter()
            po.cell_step_torch(h, params, "gs3d")
            t1 = time.perf_counter() - t0
            if t1 * 2.2 > target or d * 2 > D:
                break
            d *= 2
        for _ in range(warmup):
            h = po.cell_step_torch(h, params, "gs3d")
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            h = po.cell_step_torch(h, params, "gs3d")
            times.append(time.perf_counter() - t0)
    dt = sum(times)
    cells = d * H * W
    return {"cell_steps_per_s": cells * steps / dt, "cores": cores, "slab": (d, H, W), "seconds": dt,
            "ms_per_step": 1e3 * dt / steps,
            "sample": f"{steps} step(s) of the reference ATen op sequence (oracle port, torch {torch.__version__}, "
                      f"{cores} threads) on a {d}x{H}x{W} periodic slab of the {D}x{H}x{W} grid, scaled by cells"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    shape = wl["shape"]
    ncell = shape[0] * shape[1] * shape[2]
    r = cpu_reference_rate(shape, budget_s=max(20.0, 6.0 * (args.steps + args.warmup)), steps=args.steps, warmup=args.warmup)
    value = r["cell_steps_per_s"] / ncell
    line = {
        "impl": "reference", "metric": "timesteps/sec", "value": value, "unit": "timesteps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"], "grid": list(shape), "rollout_steps": ROLLOUT_STEPS},
        "cell_steps_per_sec": r["cell_steps_per_s"],
        "cpu_baseline": {"value": value, "unit": "timesteps/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": value, "unit": "timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    from percnn_b200 import engine
    from percnn_b200.variants import gs3d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: percnn_b200 has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = WORKLOADS[args.workload]
    shape = wl["shape"]
    D, H, W = shape
    ncell = D * H * W
    params = load_gs3d_weights()
    cell = gs3d.RCNNCell(2, 2, 5)
    cell.load_state_dict(params, strict=True)
    cell = cell.to(dev)
    flat = engine.pack_params(cell._packed_tensors(), torch.float32)
    peak, peak_src = load_peaks()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    extra = {}
    if world == 1:
        plan = engine.get_plan(cell._spec(), shape, dev)
        plan.params_load(flat)
        a = synthetic_state(shape, 0, D, dev, torch.float32)
        b = torch.empty_like(a)
        launches0 = plan.launch_count

        def rollout():
            plan.rollout_fwd(a, ROLLOUT_STEPS, h_final=b)

        for _ in range(args.warmup):
            rollout()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        with ClockSampler(local_rank) as clk:
            l0 = plan.launch_count
            ev[0].record()
            for i in range(args.steps):
                rollout()
                ev[i + 1].record()
            barrier()
            launches = plan.launch_count - l0
        step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        total_ms = ev[0].elapsed_time(ev[-1])
        assert torch.isfinite(b).all(), "rollout produced non-finite values"
        uses_tma = plan.uses_tma
        # ---- end-to-end: host buffers in, host buffer out, through percnn_rollout_fwd_host -------
        flat_h = flat.cpu().pin_memory()
        h0_h = a.cpu().pin_memory()
        e2e_reps = max(1, min(args.steps, 3))
        plan.rollout_fwd_host(flat_h, h0_h, ROLLOUT_STEPS, want_final=True)  # warm-up (allocates the staging)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(e2e_reps):
            _, fin = plan.rollout_fwd_host(flat_h, h0_h, ROLLOUT_STEPS, want_final=True)
        e2e_s = (time.perf_counter() - t0) / e2e_reps
        assert torch.equal(fin, b.cpu()), "host path and device path disagree"
        e2e = {"value": ROLLOUT_STEPS / e2e_s, "unit": "timesteps/s",
               "h2d_bytes_per_step": int(h0_h.numel() * 4 + flat_h.numel() * 4), "d2h_bytes_per_step": int(fin.numel() * 4),
               "api": "percnn_rollout_fwd_host (pinned host buffers; H2D + 500 steps + D2H per bench step)"}
        # ---- secondary: cfg4 (128^3, L2-resident) in the same run ---------------------------------
        if args.workload == "gs3d_512":
            try:
                s4 = (128, 128, 128)
                p4 = engine.get_plan(cell._spec(), s4, dev)
                p4.params_load(flat)
                a4 = synthetic_state(s4, 0, 128, dev, torch.float32)
                b4 = torch.empty_like(a4)
                for _ in range(3):
                    p4.rollout_fwd(a4, ROLLOUT_STEPS, h_final=b4)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    p4.rollout_fwd(a4, ROLLOUT_STEPS, h_final=b4)
                e1.record()
                torch.cuda.synchronize(dev)
                ms4 = e0.elapsed_time(e1) / 5
                extra["cfg4_gs3d_128"] = {"timesteps_per_s": ROLLOUT_STEPS / (ms4 * 1e-3), "ms_per_rollout": ms4,
                                          "effective_GBps": 128 ** 3 * BYTES_PER_CELL_STEP * ROLLOUT_STEPS / (ms4 * 1e-3) / 1e9,
                                          "note": "33.5 MB per step: L2-resident and launch-bound, not an HBM number"}
            except Exception as e:  # secondary measurement must never break the headline line
                extra["cfg4_gs3d_128"] = {"error": str(e)[:200]}
            # ---- secondary: one training step at 512^3 (cfg5's grid on one GPU): taped forward, fused data loss
            # (GS3D:403 pattern: every 4th state, ::2 in space), hand-derived adjoint with the loss gradient injected
            try:
                T = 8
                tape = torch.empty((T + 1, *plan.buffer_shape), dtype=torch.float32, device=dev)
                sel = tuple((t % 4 == 0) and t < T for t in range(T + 1))
                spec = engine.DataLossSpec(sel=sel, stride=2)
                tgt = torch.rand((spec.nsel, *plan.lowres_shape(2)), device=dev)

                def train_step():
                    plan.rollout_fwd(a, T, tape=tape)
                    loss = plan.data_loss_fwd(tape, T, spec, tgt)
                    return loss, plan.rollout_bwd_loss(flat, tape, T, spec, tgt)

                train_step()
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    loss, (g_h0, g_flat) = train_step()
                e1.record()
                torch.cuda.synchronize(dev)
                ms_t = e0.elapsed_time(e1) / 3 / T
                assert torch.isfinite(g_flat).all() and torch.isfinite(loss)
                extra["train_gs3d_512"] = {
                    "ms_per_timestep_fwd_plus_adjoint": ms_t, "timesteps_per_s": 1e3 / ms_t, "tape_steps": T,
                    "achieved_GBps": ncell * 40 / (ms_t * 1e-3) / 1e9, "frac_of_peak": ncell * 40 / (ms_t * 1e-3) / 1e9 / peak,
                    "note": "40 B/cell algorithmic (16 fwd + 24 adjoint, SURVEY 8d); fused data loss on states 0 and 4, stride 2; "
                            "the reference's autograd needs 146 B/cell/step of saved activations and cannot hold this grid"}
                del tape, tgt, g_h0
            except Exception as e:
                extra["train_gs3d_512"] = {"error": str(e)[:200]}
    else:
        from percnn_b200 import halo
        slab = halo.SlabRollout(cell, shape, dev, rank, world)
        slab.set_state(synthetic_state(shape, slab.z0, slab.nz, dev, torch.float32))
        for _ in range(args.warmup):
            slab.run(ROLLOUT_STEPS)
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        with ClockSampler(local_rank) as clk:
            l0 = slab.launch_count
            ev[0].record()
            for i in range(args.steps):
                slab.run(ROLLOUT_STEPS)
                ev[i + 1].record()
            barrier()
            launches = slab.launch_count - l0
        step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        total_ms = ev[0].elapsed_time(ev[-1])
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        assert torch.isfinite(slab.interior()).all()
        uses_tma = slab.plan.uses_tma
        # end-to-end: each rank uploads its slab from pinned host memory and downloads the final slab
        h_h = slab.interior().cpu().pin_memory()
        out_h = torch.empty_like(h_h).pin_memory()
        barrier()
        t0 = time.perf_counter()
        slab.set_state(h_h.to(dev, non_blocking=True))
        slab.run(ROLLOUT_STEPS)
        out_h.copy_(slab.interior(), non_blocking=True)
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": ROLLOUT_STEPS / float(t.item()), "unit": "timesteps/s", "h2d_bytes_per_step": int(h_h.numel() * 4) * world,
               "d2h_bytes_per_step": int(out_h.numel() * 4) * world, "api": "percnn_b200.halo.SlabRollout (per-rank pinned slabs)"}
        extra["halo"] = slab.describe()

    nsteps_total = ROLLOUT_STEPS * args.steps
    seconds = total_ms * 1e-3
    value = nsteps_total / seconds
    per_kernel_s = seconds / nsteps_total
    achieved = (ncell / world) * BYTES_PER_CELL_STEP / per_kernel_s / 1e9   # per GPU, per launch
    line = {
        "metric": "timesteps/sec", "value": value, "unit": "timesteps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"], "grid": list(shape), "rollout_steps": ROLLOUT_STEPS,
                   "bench_step": "one 500-timestep rollout", "parallelism": f"slab{world}" if world > 1 else "single",
                   "l2_policy": "working set (2 x state = %.0f MB per GPU) larger than L2; no flush needed" % (2 * ncell * 8 / world / 1e6)
                   if 2 * ncell * 8 / world > 126e6 else "working set fits L2 (launch-bound regime)",
                   "weights": "tests/golden/weights_gs3d.npz (reference 3d_gs_rd checkpoint)", "tma_kernel": bool(uses_tma)},
        "cell_steps_per_sec": value * ncell,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "kernel": "k_gs3d_fwd_tma",
                     "how": "16 B/cell x cells per GPU / (CUDA-event time of the rollout / launches)"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary(),
        "step_ms": step_ms,
    }
    line.update(extra)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_rate(shape, budget_s=15.0, steps=1, warmup=0)
            line["cpu_baseline"] = {"value": r["cell_steps_per_s"] / ncell, "unit": "timesteps/s", "cores": r["cores"],
                                    "kind": "port", "sample": r["sample"]}
        prof = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    line["roofline"]["traffic"] = json.load(f).get(args.workload)
            except Exception:
                pass
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gs3d_512", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
