#!/usr/bin/env python
"""Headline benchmark: timesteps/sec of the 3-D Gray-Scott Pi-block rollout (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload gs3d_512|gs3d_128]

A bench "step" is ONE full rollout (ROLLOUT_STEPS = 500 fused time steps, cfg4's rollout length) of the
V-GS3D cell (k=1, hc=2, fp32, shipped-checkpoint weights) on a 512^3 periodic grid (cfg5's grid; 1 GiB of
state, 2 GiB ping-pong working set > L2 so every step streams from HBM).  With N > 1 the grid is
slab-decomposed along D over N ranks (strong scaling; the ghost planes cross NVLink from inside the step kernel).

value   = timesteps/sec with the state resident in HBM (CUDA events, max over ranks)
e2e     = same metric through the host-buffer C-ABI call percnn_rollout_fwd_host (H2D of parameters and
          initial state from pinned memory + D2H of the final state inside the timed region)
roofline= algorithmic bytes (16 B/cell/step fp32) / event time per step kernel vs MEASURED_PEAKS.json
cpu_baseline / --impl reference = the oracle port of the reference's CPU PyTorch op sequence (the
          reference is Python and cannot travel to the GPU box) on a bounded sample of the same workload.

The same JSON line also carries, so that the driver's BENCH/SCALE records hold them:
  configs      every other BASELINE.json config on one GPU (cfg1 128^2 fp64 x200, cfg2 256^2 x1000, cfg3 512^2
               40-step BPTT for V-BUR1 and V-BUR3, cfg4 128^3 x500), each with the CPU port timed beside it (N = 1)
  train_gs3d_512  cfg5 as stated: taped forward + fused data loss (GS3D:403 pattern) + hand-derived adjoint at
               512^3 on N GPUs (40 B/cell algorithmic), per-timestep time and fraction of the HBM peak
  cfg4_gs3d_128   cfg4 (128^3 x 500) on N GPUs
  halo_check   N > 1: slab-decomposed forward AND training step vs a single-GPU recompute of the same global
               field on every rank, bit for bit (dL/dh0, states) / to 1e-6 (parameter sums)
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
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_weights(alias):
    import numpy as np
    import torch
    z = np.load(os.path.join(GOLDEN, f"weights_{alias}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def load_gs3d_weights():
    return load_weights("gs3d")


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


def smooth_state_2d(n, device, dtype, seed, lo, hi):
    """Smooth periodic 2-field state in [lo, hi] (a few Fourier modes + 1 % noise)."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    ax = torch.arange(n, dtype=torch.float64) * (2 * math.pi / n)
    x, y = torch.meshgrid(ax, ax, indexing="ij")
    fields = []
    for _ in range(2):
        f = torch.zeros(n, n, dtype=torch.float64)
        for kx in range(3):
            for ky in range(3):
                a, p1, p2 = torch.rand(3, generator=g, dtype=torch.float64)
                f += (a - 0.5) * torch.sin(kx * x + 2 * math.pi * p1) * torch.cos(ky * y + 2 * math.pi * p2)
        f = (f - f.min()) / (f.max() - f.min())
        fields.append(lo + (hi - lo) * f + 0.01 * (hi - lo) * torch.rand(n, n, generator=g, dtype=torch.float64))
    return torch.stack(fields)[None].to(dtype).to(device)


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


def cpu_config_baselines():
    """The reference's CPU PyTorch op sequence (oracle port) on the other BASELINE.json configs, bounded samples.
    Returns {cfg: {"value": timesteps/s, "unit", "cores", "kind", "sample"}}."""
    import torch
    from oracle import percnn_oracle as po
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {}

    def fwd(name, variant, params, h0, nsample, nfull):
        with torch.no_grad():
            h = po.cell_step_torch(h0, params, variant)      # warm-up
            t0 = time.perf_counter()
            for _ in range(nsample):
                h = po.cell_step_torch(h, params, variant)
            dt = time.perf_counter() - t0
        out[name] = {"value": nsample / dt, "unit": "timesteps/s", "cores": cores, "kind": "port",
                     "sample": f"{nsample} of {nfull} forward steps, full grid {tuple(h0.shape[2:])}, {h0.dtype}"}

    def bptt(name, variant, params, h0, nsample, nfull):
        p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "laplace" not in k.lower() and "filter" not in k else v)
             for k, v in params.items()}
        h = h0.clone().requires_grad_(True)
        t0 = time.perf_counter()
        outs, _ = po.rollout_torch(h, p, variant, nsample, range(nsample))
        loss = torch.cat(outs, 0)[0:-1:5, :, ::2, ::2].pow(2).mean()
        loss.backward()
        dt = time.perf_counter() - t0
        out[name] = {"value": nsample / dt, "unit": "timesteps/s (forward + autograd backward)", "cores": cores, "kind": "port",
                     "sample": f"back-propagation through {nsample} of {nfull} steps, full grid {tuple(h0.shape[2:])}, {h0.dtype}"}

    fwd("cfg1", "fwd", load_weights("fwd"), po.ic_spiral_2d(128), 200, 200)
    fwd("cfg2", "gs2d", load_weights("gs2d"), po.ic_gs_2d(256, seed=0), 200, 1000)
    bptt("cfg3i_train", "bur1", load_weights("bur1"), po.ic_fourier_2d(512, seed=1), 10, 40)
    bptt("cfg3ii_train", "bur3", po.make_phys_params("bur3"), po.ic_fourier_2d(512, seed=1, dtype=torch.float64), 10, 40)
    fwd("cfg4", "gs3d", load_gs3d_weights(), po.ic_gs_3d((128, 128, 128), seed=0), 5, 500)
    try:   # the stock modules of GS3D:41-56 on the host cores, forward + autograd backward
        torch.manual_seed(0)
        ct = torch.nn.ConvTranspose3d
        net = torch.nn.Sequential(ct(2, 8, 5, padding=2, stride=2, output_padding=1), torch.nn.Sigmoid(), ct(8, 8, 5, padding=2, stride=1),
                                  torch.nn.Conv3d(8, 2, 1))
        sd = {"convnet.%d.%s" % (i, w): getattr(net[i], w).detach() for i in (0, 2, 3) for w in ("weight", "bias")}
        low = torch.rand((1, 2, 24, 24, 24))
        prm = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        po.upscaler_torch(low, prm, "gs3d").sum().backward()       # warm-up
        t0 = time.perf_counter()
        for _ in range(3):
            po.upscaler_torch(low, prm, "gs3d").sum().backward()
        dt = (time.perf_counter() - t0) / 3
        out["upscaler_gs3d"] = {"value": dt * 1e6, "unit": "us per forward + autograd backward", "cores": cores, "kind": "port",
                                "sample": "3 calls, 24^3 -> 48^3, torch.float32"}
    except Exception as e:  # noqa: BLE001
        out["upscaler_gs3d"] = {"error": str(e)[:200]}
    return out


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

def _time_ms(fn, reps, dev):
    import torch
    fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps


def gpu_config_blocks(dev, peak):
    """cfg1-cfg4 of BASELINE.json on ONE GPU (device-resident inputs, CUDA events around whole rollouts)."""
    import torch
    from percnn_b200.variants import burgers_stage1, burgers_stage3, gs2d, gs3d, lambda_omega_fwd
    out = {}
    kw = dict(input_channels=2, hidden_channels=4, output_channels=2, input_kernel_size=5, input_stride=1, input_padding=2)

    def fwd(name, cell, h0, nsteps, desc):
        cell = cell.to(dev)
        emit = [False] * nsteps
        with torch.no_grad():
            ms = _time_ms(lambda: cell.rollout_emit(h0, nsteps, emit, want_final=True), 5, dev)
        ncell = h0[0, 0].numel()
        esz = h0.element_size()
        out[name] = {"workload": desc, "timesteps_per_s": nsteps / (ms * 1e-3), "us_per_timestep": 1e3 * ms / nsteps,
                     "ms_per_rollout": ms, "effective_GBps": ncell * 4 * esz * nsteps / (ms * 1e-3) / 1e9}

    def train(name, cell, h0, nsteps, desc):
        cell = cell.to(dev)
        h0 = h0.clone().requires_grad_(True)
        sel = [(s % 5 == 0) and s < nsteps for s in range(nsteps + 1)]
        n = h0.shape[-1]
        tgt = smooth_state_2d(n, dev, h0.dtype, 2, -0.5, 0.5)[:, :, ::2, ::2].expand(sum(sel), -1, -1, -1).contiguous()

        def step():
            for p in cell.parameters():
                p.grad = None
            h0.grad = None
            _, loss = cell.rollout_data_loss(h0, nsteps, tgt, sel, 2)
            loss.backward()

        ms = _time_ms(step, 3, dev)
        out[name] = {"workload": desc, "timesteps_per_s": nsteps / (ms * 1e-3), "us_per_timestep_fwd_plus_adjoint": 1e3 * ms / nsteps,
                     "ms_per_training_step": ms}

    def guarded(f, name, *a):
        try:
            f(name, *a)
        except Exception as e:  # a secondary measurement must never break the headline line
            out[name] = {"error": str(e)[:200]}

    c1 = lambda_omega_fwd.RCNNCell(input_kernel_size=1, input_stride=1, input_padding=0)
    c1.load_state_dict(load_weights("fwd"))
    guarded(fwd, "cfg1", c1, smooth_state_2d(128, dev, torch.float64, 1, -0.8, 0.8), 200,
            "2-D lambda-omega 128^2, fp64, hc=4, 200-step forward rollout")
    c2 = gs2d.RCNNCell(2, 8, 5)
    c2.load_state_dict(load_weights("gs2d"))
    guarded(fwd, "cfg2", c2, smooth_state_2d(256, dev, torch.float32, 1, 0.1, 0.9), 1000,
            "2-D Gray-Scott 256^2, fp32, hc=8, 1000-step forward rollout")
    c3 = burgers_stage1.RCNNCell(**kw)
    c3.load_state_dict(load_weights("bur1"))
    h3 = smooth_state_2d(512, dev, torch.float32, 1, -0.5, 0.5)
    guarded(fwd, "cfg3i_fwd", c3, h3, 40, "2-D Burgers 512^2, 5x5 Pi-block hc=16 fp32, 40-step forward")
    guarded(train, "cfg3i_train", c3, h3, 40, "2-D Burgers 512^2, 5x5 Pi-block hc=16 fp32, back-propagation through 40 steps (fused data loss)")
    c3b = burgers_stage3.RCNNCell(**kw)
    h3b = smooth_state_2d(512, dev, torch.float64, 1, -0.5, 0.5)
    guarded(fwd, "cfg3ii_fwd", c3b, h3b, 40, "2-D Burgers 512^2, d/dx d/dy advection stencil cell fp64, 40-step forward")
    guarded(train, "cfg3ii_train", c3b, h3b, 40, "2-D Burgers 512^2, advection stencil cell fp64, back-propagation through 40 steps")
    def upscale(name, mod, low_shape, desc):
        torch.manual_seed(0)
        m = mod.upscaler().to(dev)
        low = torch.rand((1, 2, *low_shape), device=dev)
        with torch.no_grad():
            g = torch.rand_like(m(low))

        def f():
            with torch.no_grad():
                m(low)

        def fb():
            m.zero_grad(set_to_none=True)
            (m(low) * g).sum().backward()

        out[name] = {"workload": desc, "forward_us": 1e3 * _time_ms(f, 10, dev), "forward_plus_adjoint_us": 1e3 * _time_ms(fb, 10, dev)}

    guarded(upscale, "upscaler_gs3d", gs3d, (24, 24, 24),
            "initial-state generator of GS3D:41-56 at the script's own size, 24^3 -> 48^3 (fused forward; forward + hand-derived adjoint)")
    c4 = gs3d.RCNNCell(2, 2, 5)
    c4.load_state_dict(load_gs3d_weights())
    guarded(fwd, "cfg4", c4, synthetic_state((128, 128, 128), 0, 128, dev, torch.float32)[None], ROLLOUT_STEPS,
            "3-D Gray-Scott 128^3, fp32, hc=2, 500-step forward rollout (L2-resident: a latency number, not an HBM one)")
    return out


def train_steps_for(world):
    """Tape length of the cfg5 training measurement: as close to the script's 150 steps as one GPU's HBM allows
    (1 GiB per stored state at N = 1)."""
    return min(150, 30 * world)


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

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    extra = {}
    T = train_steps_for(world)
    sel = tuple((t % 15 == 0) and t < T for t in range(T + 1))       # GS3D:403: output[:-1:15, :, ::2, ::2, ::2]
    train_note = ("cfg5 as stated: taped forward + fused data loss (every 15th state, ::2 in space, GS3D:403) + hand-derived "
                  "adjoint with the loss gradient injected; 40 B/cell algorithmic (16 fwd + 24 adjoint, SURVEY 8d); the "
                  "reference's autograd needs 146 B/cell/step of saved activations and cannot hold this grid")
    if world == 1:
        plan = engine.get_plan(cell._spec(), shape, dev)
        plan.params_load(flat)
        a = synthetic_state(shape, 0, D, dev, torch.float32)
        b = torch.empty_like(a)

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
        del fin, h0_h
        if args.workload == "gs3d_512" and not args.headline_only:
            # ---- cfg5 as stated, on this one GPU: one training step at 512^3 ----
            try:
                tape = torch.empty((T + 1, *plan.buffer_shape), dtype=torch.float32, device=dev)
                spec = engine.DataLossSpec(sel=sel, stride=2)
                tgt = torch.rand((spec.nsel, *plan.lowres_shape(2)), device=dev)

                def train_step():
                    plan.rollout_fwd(a, T, tape=tape)
                    loss = plan.data_loss_fwd(tape, T, spec, tgt)
                    return loss, plan.rollout_bwd_loss(flat, tape, T, spec, tgt)

                train_step()
                torch.cuda.synchronize(dev)
                reps = 2
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    loss, (g_h0, g_flat) = train_step()
                e1.record()
                torch.cuda.synchronize(dev)
                ms_t = e0.elapsed_time(e1) / reps / T
                assert torch.isfinite(g_flat).all() and torch.isfinite(loss)
                extra["train_gs3d_512"] = {
                    "ms_per_timestep_fwd_plus_adjoint": ms_t, "timesteps_per_s": 1e3 / ms_t, "tape_steps": T, "n_gpus": 1,
                    "achieved_GBps": ncell * 40 / (ms_t * 1e-3) / 1e9, "frac_of_peak": ncell * 40 / (ms_t * 1e-3) / 1e9 / peak,
                    "note": train_note}
                del tape, tgt, g_h0
            except Exception as e:
                extra["train_gs3d_512"] = {"error": str(e)[:200]}
            torch.cuda.empty_cache()
            del a, b
            engine.clear_plans()
            extra["configs"] = gpu_config_blocks(dev, peak)
            if "cfg4" in extra["configs"]:
                extra["cfg4_gs3d_128"] = dict(extra["configs"]["cfg4"], n_gpus=1)
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
        total_ms = max_over_ranks(ev[0].elapsed_time(ev[-1]))
        assert torch.isfinite(slab.interior()).all()
        uses_tma = slab.plan.uses_tma
        # end-to-end: each rank uploads its slab from pinned host memory and downloads the final slab
        h_h = slab.interior().cpu().pin_memory()
        out_h = torch.empty_like(h_h).pin_memory()

        def e2e_once():
            slab.set_state_from_host(h_h)          # two async H2D copies straight into the slab buffer + ghost exchange
            slab.run(ROLLOUT_STEPS)
            slab.interior_to_host(out_h)           # two async D2H copies
            barrier()

        e2e_once()                                 # warm-up of this path (the N = 1 leg warms its host path too)
        e2e_reps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_reps):
            e2e_once()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_reps)
        assert torch.isfinite(out_h).all()
        e2e = {"value": ROLLOUT_STEPS / e2e_s, "unit": "timesteps/s", "h2d_bytes_per_step": int(h_h.numel() * 4) * world,
               "d2h_bytes_per_step": int(out_h.numel() * 4) * world, "api": "percnn_b200.halo.SlabRollout (per-rank pinned slabs)"}
        extra["halo"] = slab.describe()
        del h_h, out_h
        if args.workload == "gs3d_512" and not args.headline_only:
            # ---- correctness of the multi-GPU path inside the timed artefact: slab vs single-GPU recompute ----
            try:
                extra["halo_check"] = halo_check(slab, cell, flat, shape, dev, world)
            except Exception as e:
                extra["halo_check"] = {"error": str(e)[:300]}
            barrier()
            # ---- cfg5 as stated: domain-decomposed training rollout on N GPUs ----
            try:
                spec_sel = sel
                nsel = sum(spec_sel)
                tgt = torch.rand((nsel, 2, slab.nz // 2, H // 2, W // 2), device=dev)
                h0_slab = synthetic_state(shape, slab.z0, slab.nz, dev, torch.float32)
                gscale = torch.tensor(10.0, device=dev)        # `10*loss_data` (GS3D:407)

                def train_step():
                    slab.set_state(h0_slab)
                    tape = slab.rollout_tape(T)
                    loss = slab.data_loss(tape, tgt, spec_sel, 2)
                    g_h0, grads = slab.backward(tape, None, loss=(tgt, spec_sel, 2, gscale))
                    return loss, grads

                train_step()
                barrier()
                reps = 2
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    loss, grads = train_step()
                e1.record()
                barrier()
                ms_t = max_over_ranks(e0.elapsed_time(e1)) / reps / T
                ok = bool(torch.isfinite(grads).all() and torch.isfinite(loss))
                extra["train_gs3d_512"] = {
                    "ms_per_timestep_fwd_plus_adjoint": ms_t, "timesteps_per_s": 1e3 / ms_t, "tape_steps": T, "n_gpus": world,
                    "achieved_GBps_per_gpu": ncell / world * 40 / (ms_t * 1e-3) / 1e9,
                    "frac_of_peak_per_gpu": ncell / world * 40 / (ms_t * 1e-3) / 1e9 / peak, "finite": ok,
                    "includes": "set_state (ghost exchange + barriers), taped forward, fused loss (1 scalar all-reduce), fused-halo adjoint, "
                                "one 24-value all-reduce of the parameter sums", "note": train_note}
                del tgt, h0_slab
                slab._tape = None
                slab._tape_shape = None
            except Exception as e:
                extra["train_gs3d_512"] = {"error": str(e)[:300]}
            barrier()
            torch.cuda.empty_cache()
            # ---- cfg4 as stated: 128^3 x 500 steps on N GPUs ----
            try:
                s4 = (128, 128, 128)
                slab4 = halo.SlabRollout(cell, s4, dev, rank, world)
                slab4.set_state(synthetic_state(s4, slab4.z0, slab4.nz, dev, torch.float32))
                for _ in range(3):
                    slab4.run(ROLLOUT_STEPS)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    slab4.run(ROLLOUT_STEPS)
                e1.record()
                barrier()
                ms4 = max_over_ranks(e0.elapsed_time(e1)) / 5
                # correctness of THIS path inside the artefact: the full 500-step slab rollout against a single-GPU
                # recompute of the same global field on every rank (the persistent slab kernel shares the gather
                # kernel's per-cell arithmetic: bitwise; the TMA z-march differs from it by rounding only)
                from percnn_b200 import _lib as _plib
                full4 = synthetic_state(s4, 0, s4[0], dev, torch.float32)
                slab4.set_state(full4[:, slab4.z0:slab4.z0 + slab4.nz])
                slab4.run(ROLLOUT_STEPS)
                ref_flags = _plib.FLAG_NO_TMA if slab4.plan.slab_persistent else 0
                import dataclasses as _dc
                plan4 = engine.get_plan(_dc.replace(cell._spec(), flags=ref_flags), s4, dev)
                plan4.params_load(flat)
                ref4 = torch.empty_like(full4)
                plan4.rollout_fwd(full4, ROLLOUT_STEPS, h_final=ref4)
                same = torch.tensor([int(torch.equal(slab4.interior(), ref4[:, slab4.z0:slab4.z0 + slab4.nz]))], device=dev)
                dist.all_reduce(same, op=dist.ReduceOp.MIN)
                extra["cfg4_gs3d_128"] = {"workload": "3-D Gray-Scott 128^3, fp32, hc=2, 500-step forward rollout, slab-decomposed",
                                          "n_gpus": world, "timesteps_per_s": ROLLOUT_STEPS / (ms4 * 1e-3), "us_per_timestep": 1e3 * ms4 / ROLLOUT_STEPS,
                                          "ms_per_rollout": ms4, "halo": slab4.describe(), "finite": bool(torch.isfinite(slab4.interior()).all()),
                                          "vs_single_gpu_500_steps": "bitwise" if int(same.item()) else "MISMATCH"}
            except Exception as e:
                extra["cfg4_gs3d_128"] = {"error": str(e)[:300]}

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
                     "traffic": None, "peak_source": peak_src, "kernel": "k_gs3d_fwd_tma" if world == 1 else "k_gs3d_fwd_slab",
                     "how": "16 B/cell x cells per GPU / (CUDA-event time of the rollout / launches)"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary(),
        "step_ms": step_ms,
    }
    line.update(extra)
    if rank == 0:
        if world == 1:
            # dram bytes of one launch from the committed `ncu --set full` capture of THIS workload on one GPU
            prof = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(prof):
                try:
                    with open(prof) as f:
                        line["roofline"]["traffic"] = json.load(f).get(args.workload)
                except Exception:
                    pass
            if not args.no_cpu_baseline:
                r = cpu_reference_rate(shape, budget_s=15.0, steps=1, warmup=0)
                line["cpu_baseline"] = {"value": r["cell_steps_per_s"] / ncell, "unit": "timesteps/s", "cores": r["cores"],
                                        "kind": "port", "sample": r["sample"]}
                if "configs" in line:
                    try:
                        for k, v in cpu_config_baselines().items():
                            line["configs"].setdefault(k, {})["cpu_baseline"] = v
                    except Exception as e:
                        line["configs"]["cpu_baseline_error"] = str(e)[:200]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def halo_check(slab, cell, flat, shape, dev, world):
    """Slab-decomposed rollout and training step vs a single-GPU recompute of the SAME global field on every rank.
    5 forward steps (both march directions, odd count) and a 4-step taped forward + fused-loss adjoint."""
    import torch
    import torch.distributed as dist
    from percnn_b200 import engine
    D, H, W = shape
    z0, nz = slab.z0, slab.nz
    full = synthetic_state(shape, 0, D, dev, torch.float32, seed=11)      # identical on every rank
    plan1 = engine.get_plan(cell._spec(), shape, dev)
    plan1.params_load(flat)
    ref = torch.empty_like(full)
    plan1.rollout_fwd(full, 5, h_final=ref)
    slab.set_state(full[:, z0:z0 + nz])
    slab.run(5)
    fwd_ok = bool(torch.equal(slab.interior(), ref[:, z0:z0 + nz]))
    del ref
    # training step: T = 4, loss on states 0 and 2, stride 2
    T = 4
    sel = (True, False, True, False, False)
    spec = engine.DataLossSpec(sel=sel, stride=2)
    g = torch.Generator(device=dev).manual_seed(5)
    tgt_full = torch.rand((2, *plan1.lowres_shape(2)), generator=g, device=dev)
    tape1 = torch.empty((T + 1, *plan1.buffer_shape), dtype=torch.float32, device=dev)
    plan1.rollout_fwd(full, T, tape=tape1)
    loss1 = plan1.data_loss_fwd(tape1, T, spec, tgt_full)
    g_h0_ref, g_flat_ref = plan1.rollout_bwd_loss(flat, tape1, T, spec, tgt_full)
    slab.set_state(full[:, z0:z0 + nz])
    tape = slab.rollout_tape(T)
    tape_ok = bool(torch.equal(tape[:, :, 2:nz + 2], tape1[:, :, z0:z0 + nz]))
    tgt_loc = tgt_full[:, :, z0 // 2:(z0 + nz) // 2].contiguous()
    loss = slab.data_loss(tape, tgt_loc, sel, 2)
    g_h0, grads = slab.backward(tape, None, loss=(tgt_loc, sel, 2, None))
    gh_ok = bool(torch.equal(g_h0, g_h0_ref[:, z0:z0 + nz]))
    gp_err = float((grads.double() - g_flat_ref.double()).norm() / g_flat_ref.double().norm())
    loss_err = abs(float(loss) - float(loss1)) / abs(float(loss1))
    slab._tape = None
    slab._tape_shape = None
    del tape, tape1, full
    flags = torch.tensor([int(fwd_ok), int(tape_ok), int(gh_ok), int(gp_err < 1e-6), int(loss_err < 1e-6)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    f = [bool(x) for x in flags.tolist()]
    res = {"forward_5_steps": "bitwise" if f[0] else "MISMATCH", "taped_forward_4_steps": "bitwise" if f[1] else "MISMATCH",
           "adjoint_dL_dh0": "bitwise" if f[2] else "MISMATCH", "param_grad_rel_err": gp_err, "param_grads_ok": f[3],
           "loss_rel_err": loss_err, "loss_ok": f[4], "ranks": world, "grid": list(shape),
           "against": "single-GPU recompute of the same global field on every rank (k_gs3d_fwd_tma / k_gs3d_bwd_tma)"}
    res["ok"] = all(f)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gs3d_512", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the secondary blocks (configs, training, halo check)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
