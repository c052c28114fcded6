"""Single GPU: the same rollout twice must be bitwise identical (catches intra-GPU ordering races)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import load_gs3d_weights, synthetic_state
from percnn_b200 import engine
from percnn_b200.variants import gs3d
dev = torch.device("cuda:0")
cell = gs3d.RCNNCell(2, 2, 5); cell.load_state_dict(load_gs3d_weights()); cell = cell.to(dev)
for shape in [(64, 512, 512), (128, 128, 128), (128, 512, 512)]:
    plan = engine.get_plan(cell._spec(), shape, dev)
    plan.params_load(engine.pack_params(cell._packed_tensors(), torch.float32))
    a = synthetic_state(shape, 0, shape[0], dev, torch.float32, seed=3)
    outs = []
    for rep in range(4):
        b = torch.empty_like(a)
        plan.rollout_fwd(a, 25, h_final=b)
        torch.cuda.synchronize()
        outs.append(b)
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    cell._flags = 2  # NO_TMA generic kernel as an independent reference
    plan2 = engine.get_plan(cell._spec(), shape, dev)
    plan2.params_load(engine.pack_params(cell._packed_tensors(), torch.float32))
    c = torch.empty_like(a); plan2.rollout_fwd(a, 25, h_final=c); torch.cuda.synchronize()
    cell._flags = 0
    print(f"DETERMINISM shape={shape} NO_PDL={os.environ.get('PERCNN_NO_PDL','0')}: repeats identical={same}  max|tma-generic|={float((outs[0]-c).abs().max()):.3e}", flush=True)
