"""A few launches of each flagship kernel for `ncu --set full` (one GPU): the periodic 3-D forward step and its adjoint at
512^3, the slab-mode step (ring of one rank, 64 planes of 512^2 = the per-rank slab of cfg5 on 8 GPUs), the 2-D tiled
rollout (cfg2) and the 5x5 Pi-block step (cfg3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import load_gs3d_weights, load_weights, smooth_state_2d, synthetic_state  # noqa: E402
from percnn_b200 import engine, halo  # noqa: E402
from percnn_b200.variants import burgers_stage1, gs2d, gs3d  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
cell = gs3d.RCNNCell(2, 2, 5)
cell.load_state_dict(load_gs3d_weights())
cell = cell.to(dev)
flat = engine.pack_params(cell._packed_tensors(), torch.float32)
if which in ("all", "fwd", "bwd"):
    shape = (512, 512, 512)
    plan = engine.get_plan(cell._spec(), shape, dev)
    plan.params_load(flat)
    a = synthetic_state(shape, 0, 512, dev, torch.float32)
    T = 3
    tape = torch.empty((T + 1, *plan.buffer_shape), device=dev)
    plan.rollout_fwd(a, T, tape=tape)
    if which in ("all", "bwd"):
        spec = engine.DataLossSpec(sel=(True, False, False, False), stride=2)
        tgt = torch.rand((1, *plan.lowres_shape(2)), device=dev)
        plan.rollout_bwd_loss(flat, tape, T, spec, tgt)
    torch.cuda.synchronize()
    del tape, a, plan
    engine.clear_plans()
    torch.cuda.empty_cache()
if which in ("all", "slab"):
    shape = (64, 512, 512)
    slab = halo.SlabRollout(cell, shape, dev, 0, 1, transport="fused")
    slab.set_state(synthetic_state(shape, 0, 64, dev, torch.float32))
    slab.run(6)
    torch.cuda.synchronize()
    del slab
    engine.clear_plans()
if which in ("all", "slabtrain"):
    # the cfg5 training step's kernels on the per-rank slab of 8 GPUs: taped slab forward + fused-halo adjoint (ring of one)
    shape = (64, 512, 512)
    slab = halo.SlabRollout(cell, shape, dev, 0, 1, transport="fused")
    slab.set_state(synthetic_state(shape, 0, 64, dev, torch.float32))
    T = 4
    tape = slab.rollout_tape(T)
    sel = (True, False, False, False, False)
    tgt = torch.rand((1, 2, 32, 256, 256), device=dev)
    slab.backward(tape, None, loss=(tgt, sel, 2, None))
    torch.cuda.synchronize()
    del slab, tape
    engine.clear_plans()
    # and the periodic adjoint on the same 64 x 512 x 512 cells, for comparison
    plan = engine.get_plan(cell._spec(), shape, dev)
    plan.params_load(flat)
    a = synthetic_state(shape, 0, 64, dev, torch.float32)
    tape = torch.empty((T + 1, *plan.buffer_shape), device=dev)
    plan.rollout_fwd(a, T, tape=tape)
    spec = engine.DataLossSpec(sel=sel, stride=2)
    plan.rollout_bwd_loss(flat, tape, T, spec, tgt[0:1].reshape(1, 2, 32, 256, 256))
    torch.cuda.synchronize()
    del tape, a, plan
    engine.clear_plans()
    torch.cuda.empty_cache()
if which in ("all", "slabsmall"):
    # the communication-avoiding persistent kernel on cfg4's per-rank slab of 8 GPUs (ring of one)
    shape = (16, 128, 128)
    slab = halo.SlabRollout(cell, shape, dev, 0, 1, transport="fused")
    slab.set_state(synthetic_state(shape, 0, 16, dev, torch.float32))
    slab.run(41)
    torch.cuda.synchronize()
    del slab
    engine.clear_plans()
if which in ("all", "tile2d"):
    c2 = gs2d.RCNNCell(2, 8, 5)
    c2.load_state_dict(load_weights("gs2d"))
    c2 = c2.to(dev)
    with torch.no_grad():
        c2.rollout_emit(smooth_state_2d(256, dev, torch.float32, 1, 0.1, 0.9), 200, [False] * 200, want_final=True)
    torch.cuda.synchronize()
if which in ("all", "k5"):
    kw = dict(input_channels=2, hidden_channels=4, output_channels=2, input_kernel_size=5, input_stride=1, input_padding=2)
    c3 = burgers_stage1.RCNNCell(**kw)
    c3.load_state_dict(load_weights("bur1"))
    c3 = c3.to(dev)
    h = smooth_state_2d(512, dev, torch.float32, 1, -0.5, 0.5).requires_grad_(True)
    c3.rollout(h, 2).square().sum().backward()
    torch.cuda.synchronize()
