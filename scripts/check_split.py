"""Single GPU: three-segment schedule (PERCNN_TMA_SPLIT=1, child process) vs the plain one-segment step, bitwise."""
import os, subprocess, sys
CHILD = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from bench import load_gs3d_weights, synthetic_state
from percnn_b200 import engine
from percnn_b200.variants import gs3d
dev = torch.device("cuda:0")
cell = gs3d.RCNNCell(2, 2, 5); cell.load_state_dict(load_gs3d_weights()); cell = cell.to(dev)
shape = (64, 512, 512)
plan = engine.get_plan(cell._spec(), shape, dev)
plan.params_load(engine.pack_params(cell._packed_tensors(), torch.float32))
a = synthetic_state(shape, 0, shape[0], dev, torch.float32, seed=3)
ref = None
nbad = 0
for rep in range(int(os.environ.get("REPS", "200"))):
    b = torch.empty_like(a)
    plan.rollout_fwd(a, 1, h_final=b)
    torch.cuda.synchronize()
    if ref is None:
        ref = b.clone(); torch.save(ref.cpu(), os.environ["OUT"])
    elif not torch.equal(ref, b):
        bad = ref != b
        nbad += 1
        if nbad <= 4:
            zs = bad.any(0).flatten(1).any(1).nonzero().flatten().tolist(); ys = bad.any(0).any(0).any(1).nonzero().flatten().tolist()
            print("  rep", rep, "cells", int(bad.sum()), "fields", bad.flatten(1).any(1).nonzero().flatten().tolist(), "planes", zs[:6], "rows", ys[:6], "maxerr", float((ref-b).abs().max()))
print("split=%s reps identical: %s (bad %d)" % (os.environ.get("PERCNN_TMA_SPLIT", "0"), nbad == 0, nbad))
'''
outs = []
for split in ("0", "1"):
    out = f"/tmp/split_{split}.pt"
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, PERCNN_TMA_SPLIT=split, OUT=out), capture_output=True, text=True, timeout=200)
    print(r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
    outs.append(out)
import torch
a, b = torch.load(outs[0]), torch.load(outs[1])
print("split vs plain bitwise equal:", torch.equal(a, b), "max diff", float((a - b).abs().max()))
