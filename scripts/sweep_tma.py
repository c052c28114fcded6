"""Experiment sweep for the TMA kernel (one process per configuration; knobs are read from the environment
at plan creation): prints ms/step and effective GB/s for each."""
import os
import subprocess
import sys

CHILD = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from bench import load_gs3d_weights, synthetic_state
from percnn_b200 import engine
from percnn_b200.variants import gs3d
n = int(os.environ.get("N", "512")); steps = int(os.environ.get("STEPS", "40"))
dev = torch.device("cuda:0")
cell = gs3d.RCNNCell(2, 2, 5); cell.load_state_dict(load_gs3d_weights()); cell = cell.to(dev)
shape = (n, n, n)
plan = engine.get_plan(cell._spec(), shape, dev)
plan.params_load(engine.pack_params(cell._packed_tensors(), torch.float32))
a = synthetic_state(shape, 0, n, dev, torch.float32); b = torch.empty_like(a)
plan.rollout_fwd(a, 10, h_final=b); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record(); plan.rollout_fwd(a, steps, h_final=b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / steps)
print(f"{best*1e3:9.1f} us/step  {n**3*16/best/1e6:8.1f} GB/s")
if os.environ.get("COPYREF"):
    x = torch.empty(1 << 29, dtype=torch.float32, device=dev); y = torch.empty_like(x)
    y.copy_(x); torch.cuda.synchronize()
    e0.record()
    for _ in range(10): y.copy_(x)
    e1.record(); torch.cuda.synchronize()
    print(f"torch copy 2 GiB: {x.numel()*8*10/e0.elapsed_time(e1)/1e6:8.1f} GB/s")
'''

configs = [
    ("default+copyref", {"COPYREF": "1"}),
    ("ty=16", {"PERCNN_TMA_TY": "16"}),
    ("ty=14", {"PERCNN_TMA_TY": "14"}),
    ("ty=13", {"PERCNN_TMA_TY": "13"}),
    ("ty=7", {"PERCNN_TMA_TY": "7"}),
    ("default skeleton", {"PERCNN_TMA_MODE": "1"}),
    ("default streaming stores", {"PERCNN_TMA_MODE": "4"}),
    ("default tz=128", {"PERCNN_TMA_TZ": "128"}),
    ("N=256", {"N": "256", "STEPS": "200"}),
    ("N=256 ty=16", {"N": "256", "STEPS": "200", "PERCNN_TMA_TY": "16"}),
    ("N=256 ty=8", {"N": "256", "STEPS": "200", "PERCNN_TMA_TY": "8"}),
    ("N=128", {"N": "128", "STEPS": "500"}),
    ("N=128 ty=16", {"N": "128", "STEPS": "500", "PERCNN_TMA_TY": "16"}),
    ("N=128 ty=8", {"N": "128", "STEPS": "500", "PERCNN_TMA_TY": "8"}),
    ("N=128 ty=4", {"N": "128", "STEPS": "500", "PERCNN_TMA_TY": "4"}),
]
for name, env in configs:
    e = dict(os.environ)
    e.update(env)
    try:
        out = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=120)
        print(f"{name:28s} {out.stdout.strip() or out.stderr.strip()[-300:]}", flush=True)
    except subprocess.TimeoutExpired:
        print(f"{name:28s} TIMEOUT", flush=True)
