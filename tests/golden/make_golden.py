"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

The reference has no tests or golden vectors for its hot path, so parity is
pinned here: this script imports the reference scripts' own `RCNNCell` / `RCNN` classes from
/root/reference (read-only, never copied), loads the weights of the shipped checkpoints, drives
the classes with small seeded states and records inputs, parameters, outputs and autograd
gradients.  /root/reference does not exist on the GPU box, so the vectors are committed.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

Shims needed to import the scripts: stub matplotlib/plotly/prettytable,
importlib by path (file names contain '[' ',' '%'), no-op .cuda(), snapshot/restore of the
default dtype and CUDA_VISIBLE_DEVICES that the scripts set at import time.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

SCRIPTS = {
    "fwd": "ForwardSimulationOfPDEs/2d_lambda_omega/percnn_LO_eqn.py",
    "gs2d": "DataDrivenModeling/2d_gs_rd/train_2drd.py",
    "gs3d": "DataDrivenModeling/3d_gs_rd/train_3drd.py",
    "bur1": "DataDrivenDiscoveryOfPDEs/2D_Burgers_eqn/Stage-1/rcnn_Burgers_[resnet,GT41x51x51,LAPLACE,5%noise].py",
    "lo1": "DataDrivenDiscoveryOfPDEs/2D_Lambda_Omega_eqn/stage-1/rcnn_LO_[resnet,GT41x51x51,LAPLACE,5%noise].py",
    "bur3": "DataDrivenDiscoveryOfPDEs/2D_Burgers_eqn/Stage-3/fine_tuning_[5%noise,41x51x51].py",
    "lo3": "DataDrivenDiscoveryOfPDEs/2D_Lambda_Omega_eqn/stage-3/fine_tuning_LO_[0%noise,41x51x51].py",
    "lo3n": "DataDrivenDiscoveryOfPDEs/2D_Lambda_Omega_eqn/stage-3/fine_tuning_LO_[10%noise,41x51x51].py",
}
CHECKPOINTS = {
    "fwd": "ForwardSimulationOfPDEs/2d_lambda_omega/model/rcnn_pde.pt",
    "gs2d": "DataDrivenModeling/2d_gs_rd/model/checkpoint.pt",
    "gs3d": "DataDrivenModeling/3d_gs_rd/model/checkpoint.pt",
    "bur1": "DataDrivenDiscoveryOfPDEs/2D_Burgers_eqn/Stage-1/model/checkpoint.pt",
    "lo1": "DataDrivenDiscoveryOfPDEs/2D_Lambda_Omega_eqn/stage-1/model/checkpoint.pt",
}


def load_reference_module(alias):
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "plotly", "plotly.graph_objects",
                 "prettytable", "mpl_toolkits", "mpl_toolkits.axes_grid1", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["mpl_toolkits.axes_grid1"].make_axes_locatable = lambda *a, **k: None
    sys.modules["matplotlib"].gridspec = sys.modules["matplotlib.gridspec"]
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    env = os.environ.get("CUDA_VISIBLE_DEVICES")
    dt = torch.get_default_dtype()
    spec = importlib.util.spec_from_file_location("percnn_ref_" + alias, os.path.join(REF, SCRIPTS[alias]))
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    finally:
        mod_dtype = torch.get_default_dtype()
        torch.set_default_dtype(dt)
        if env is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = env
    mod._default_dtype = mod_dtype
    return mod


def cell_state(cell):
    return {k: v.detach().clone() for k, v in cell.state_dict().items()}


def build_cell(alias, mod):
    torch.set_default_dtype(mod._default_dtype)
    try:
        if alias == "fwd":
            cell = mod.RCNNCell(input_kernel_size=1, input_stride=1, input_padding=0)
        elif alias == "gs2d":
            cell = mod.RCNNCell(input_channels=2, hidden_channels=8, input_kernel_size=5)
        elif alias == "gs3d":
            cell = mod.RCNNCell(input_channels=2, hidden_channels=2, input_kernel_size=5)
        else:
            cell = mod.RCNNCell(input_channels=2, hidden_channels=4, output_channels=2,
                                input_kernel_size=5, input_stride=1, input_padding=2)
    finally:
        torch.set_default_dtype(torch.float32)
    if alias in CHECKPOINTS:
        ck = torch.load(os.path.join(REF, CHECKPOINTS[alias]), map_location="cpu", weights_only=False)
        sd = ck["model_state_dict"] if "model_state_dict" in ck else ck
        sd = {k.split("cell.", 1)[1]: v for k, v in sd.items() if "cell." in k}
        cell.load_state_dict(sd, strict=True)
    return cell


def smooth_state(shape, seed, dtype, lo=0.1, hi=0.9):
    """Smooth periodic field in [lo, hi] plus a little noise; small enough to commit."""
    g = torch.Generator().manual_seed(seed)
    nd = len(shape)
    grids = torch.meshgrid(*[torch.arange(n, dtype=torch.float64) * (2 * np.pi / n) for n in shape], indexing="ij")
    fields = []
    for f in range(2):
        a = torch.zeros(shape, dtype=torch.float64)
        for _ in range(4):
            ks = torch.randint(0, 3, (nd,), generator=g)
            ph = torch.rand(nd, generator=g, dtype=torch.float64) * 2 * np.pi
            term = torch.ones(shape, dtype=torch.float64)
            for d in range(nd):
                term = term * torch.cos(ks[d] * grids[d] + ph[d])
            a = a + term * (torch.rand((), generator=g, dtype=torch.float64) + 0.2)
        a = (a - a.min()) / (a.max() - a.min())
        fields.append(lo + (hi - lo) * a + 0.01 * torch.randn(shape, generator=g, dtype=torch.float64))
    return torch.stack(fields)[None].to(dtype)


def reference_rollout(cell, h0, step, effective_step):
    """The loop of RCNN.forward (GS2D:169-188) driving the reference's own cell."""
    h = h0
    outputs = [h0]
    second_last = None
    for s in range(step):
        h, o = cell(h)
        if s == step - 2:
            second_last = h.clone()
        if s in effective_step:
            outputs.append(o)
    return outputs, second_last


def make_case(alias, shape, nstep, seed, tag=None, amp=(0.1, 0.9)):
    mod = load_reference_module(alias)
    cell = build_cell(alias, mod)
    dtype = mod._default_dtype
    if alias in ("bur3", "lo3", "lo3n"):
        # jitter the literals so that every coefficient gradient is exercised off its initial value
        g = torch.Generator().manual_seed(seed + 100)
        with torch.no_grad():
            for n, prm in cell.named_parameters():
                if prm.requires_grad:
                    prm.add_(0.05 * (torch.rand((), generator=g, dtype=torch.float64) - 0.5))
    h0 = smooth_state(shape, seed, dtype, *amp).requires_grad_(True)
    eff = list(range(nstep))
    outputs, second_last = reference_rollout(cell, h0, nstep, eff)
    traj = torch.cat(outputs, dim=0)  # [nstep+1, 2, ...]
    g = torch.Generator().manual_seed(seed + 7)
    wts = torch.randn(traj.shape, generator=g, dtype=torch.float64).to(dtype)
    wts[0] = 0  # the loss never sees the given initial frame except through the dynamics
    loss = (traj * wts).sum()
    loss.backward()
    rec = {"h0": h0.detach().numpy(), "traj": traj.detach().numpy(), "second_last": second_last.detach().numpy(),
           "loss_weights": wts.numpy(), "loss": np.array(loss.item()), "g_h0": h0.grad.numpy(),
           "nstep": np.array(nstep), "dtype": np.array(str(dtype))}
    for k, v in cell_state(cell).items():
        rec["param/" + k] = v.numpy()
    for n, prm in cell.named_parameters():
        if prm.requires_grad:
            rec["grad/" + n] = prm.grad.numpy()
    name = f"cell_{tag or alias}.npz"
    np.savez_compressed(os.path.join(OUT, name), **rec)
    print(f"{name}: shape={tuple(h0.shape)} steps={nstep} dtype={dtype} loss={loss.item():.6g} "
          f"{os.path.getsize(os.path.join(OUT, name)) / 1024:.0f} KiB")


def make_data_loss_case(alias, shape, nstep, seed, t_stride, s_stride, first_frames=None, tag=None, amp=(0.1, 0.9)):
    """The scripts' data loss through the reference's own cell, loop and nn.MSELoss (GS3D:394-403, GS2D:394-401,
    BUR1:606-614): loss = mse_loss(cat(outputs)[0:-1:t, :, ::s, ::s(, ::s)][:first_frames], truth_sub), and its
    autograd gradients w.r.t. the initial state and every trainable parameter."""
    mod = load_reference_module(alias)
    cell = build_cell(alias, mod)
    dtype = mod._default_dtype
    if alias in ("bur3", "lo3", "lo3n"):
        g = torch.Generator().manual_seed(seed + 100)
        with torch.no_grad():
            for n, prm in cell.named_parameters():
                if prm.requires_grad:
                    prm.add_(0.05 * (torch.rand((), generator=g, dtype=torch.float64) - 0.5))
    h0 = smooth_state(shape, seed, dtype, *amp).requires_grad_(True)
    outputs, _ = reference_rollout(cell, h0, nstep, list(range(nstep)))
    output = torch.cat(tuple(outputs), dim=0)
    sub = (slice(None), slice(None)) + (slice(None, None, s_stride),) * len(shape)
    pred = output[0:-1:t_stride][sub]
    if first_frames is not None:
        pred = pred[:first_frames]
    g = torch.Generator().manual_seed(seed + 9)
    truth_sub = (pred.detach().double() + 0.05 * torch.randn(pred.shape, generator=g, dtype=torch.float64)).to(dtype)
    loss = torch.nn.MSELoss()(pred, truth_sub)
    (3.0 * loss).backward()           # a non-unit upstream gradient, like `10*loss_data` in GS3D:407
    frames = list(range(0, nstep, t_stride))[:first_frames]
    rec = {"h0": h0.detach().numpy(), "traj": output.detach().numpy(), "truth_sub": truth_sub.numpy(),
           "loss": np.array(loss.item()), "gscale": np.array(3.0), "g_h0": h0.grad.numpy(), "nstep": np.array(nstep),
           "t_stride": np.array(t_stride), "s_stride": np.array(s_stride), "frames": np.array(frames),
           "first_frames": np.array(-1 if first_frames is None else first_frames), "dtype": np.array(str(dtype))}
    for k, v in cell_state(cell).items():
        rec["param/" + k] = v.numpy()
    for n, prm in cell.named_parameters():
        if prm.requires_grad:
            rec["grad/" + n] = prm.grad.numpy()
    name = f"dloss_{tag or alias}.npz"
    np.savez_compressed(os.path.join(OUT, name), **rec)
    print(f"{name}: shape={tuple(h0.shape)} steps={nstep} frames={frames} stride={s_stride} loss={loss.item():.6g} "
          f"{os.path.getsize(os.path.join(OUT, name)) / 1024:.0f} KiB")


def make_data_loss_cases():
    make_data_loss_case("gs2d", (22, 26), 9, 21, 4, 4, first_frames=2)
    make_data_loss_case("gs3d", (6, 8, 12), 7, 22, 3, 2)
    make_data_loss_case("gs3d", (8, 16, 128), 4, 23, 2, 2, tag="gs3d_tma")
    make_data_loss_case("bur1", (16, 20), 6, 24, 5, 2, amp=(-0.5, 0.5))
    make_data_loss_case("bur3", (20, 24), 6, 25, 5, 2, amp=(-0.5, 0.5))
    make_data_loss_case("lo3", (20, 24), 6, 26, 5, 2, amp=(-0.8, 0.8))


def make_phys_case(alias, shape, nstep, seed, amp=(0.1, 0.9)):
    """The scripts' physics-residual loss through the reference's own `loss_generator` + `loss_gen`/`loss_func`
    (FWD:265-357, GS2D:241-353, GS3D:264-345) on a trajectory produced by the reference's own cell: the loss, its
    gradient w.r.t. the trajectory, and the end-to-end gradients w.r.t. h0 and the cell parameters."""
    mod = load_reference_module(alias)
    cell = build_cell(alias, mod)
    dtype = mod._default_dtype
    # the shipped weights solve their PDE (rcnn_pde.pt gives a residual of 1e-15): jitter them so that the loss
    # and every gradient are exercised away from zero
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for n, prm in cell.named_parameters():
            if prm.requires_grad:
                prm.mul_(1.0 + 0.2 * (torch.rand(prm.shape, generator=g, dtype=torch.float64).to(prm.dtype) - 0.5))
    h0 = smooth_state(shape, seed, dtype, *amp).requires_grad_(True)
    outputs, _ = reference_rollout(cell, h0, nstep, list(range(nstep)))
    output = torch.cat(tuple(outputs), dim=0)
    torch.set_default_dtype(dtype)
    try:
        gen = mod.loss_generator()      # script defaults: dt, dx of that PDE
        fn = mod.loss_func if alias == "gs3d" else mod.loss_gen
        leaf = output.detach().clone().requires_grad_(True)
        loss_leaf = fn(leaf, gen)
        loss_leaf.backward()
        loss = fn(output, gen)
        (2.0 * loss).backward()
    finally:
        torch.set_default_dtype(torch.float32)
    rec = {"h0": h0.detach().numpy(), "traj": output.detach().numpy(), "loss": np.array(loss.item()),
           "g_traj": leaf.grad.numpy(), "gscale": np.array(2.0), "g_h0": h0.grad.numpy(), "nstep": np.array(nstep),
           "dtype": np.array(str(dtype))}
    for k, v in cell_state(cell).items():
        rec["param/" + k] = v.numpy()
    for n, prm in cell.named_parameters():
        if prm.requires_grad:
            rec["grad/" + n] = prm.grad.numpy()
    name = f"phys_{alias}.npz"
    np.savez_compressed(os.path.join(OUT, name), **rec)
    print(f"{name}: shape={tuple(h0.shape)} steps={nstep} dtype={dtype} loss={loss.item():.6g} "
          f"{os.path.getsize(os.path.join(OUT, name)) / 1024:.0f} KiB")


def make_phys_cases():
    make_phys_case("fwd", (18, 22), 5, 31, amp=(-0.8, 0.8))
    make_phys_case("gs2d", (18, 22), 5, 32)
    make_phys_case("gs3d", (6, 7, 9), 4, 33)


def make_rk4_cases():
    """`RCNNCell.forward_rk4` of the Stage-3 scripts (BUR3:159-206; defined, never called by the scripts) driven for
    three steps, with autograd gradients of a weighted sum w.r.t. the initial state and every coefficient."""
    for alias, seed, amp in (("bur3", 41, (-0.5, 0.5)), ("lo3", 42, (-0.8, 0.8)), ("lo3n", 43, (-0.8, 0.8))):
        mod = load_reference_module(alias)
        cell = build_cell(alias, mod)
        g = torch.Generator().manual_seed(seed + 100)
        with torch.no_grad():
            for n, prm in cell.named_parameters():
                if prm.requires_grad:
                    prm.add_(0.05 * (torch.rand((), generator=g, dtype=torch.float64) - 0.5))
        h0 = smooth_state((20, 24), seed, torch.float64, *amp).requires_grad_(True)
        h, traj = h0, [h0]
        for _ in range(3):
            h, _ = cell.forward_rk4(h)
            traj.append(h)
        traj = torch.cat(traj, 0)
        w = torch.randn(traj.shape, generator=torch.Generator().manual_seed(seed + 7), dtype=torch.float64)
        loss = (traj * w).sum()
        loss.backward()
        rec = {"h0": h0.detach().numpy(), "traj": traj.detach().numpy(), "loss_weights": w.numpy(),
               "loss": np.array(loss.item()), "g_h0": h0.grad.numpy()}
        for k, v in cell.state_dict().items():
            rec["param/" + k] = v.detach().numpy()
        for n, prm in cell.named_parameters():
            if prm.requires_grad:
                rec["grad/" + n] = prm.grad.numpy()
        np.savez_compressed(os.path.join(OUT, f"rk4_{alias}.npz"), **rec)
        print(f"rk4_{alias}.npz", tuple(traj.shape), loss.item())


def make_weights():
    """Cell weights of the shipped checkpoints at full size (bench + full-size property tests)."""
    for alias in CHECKPOINTS:
        mod = load_reference_module(alias)
        cell = build_cell(alias, mod)
        np.savez_compressed(os.path.join(OUT, f"weights_{alias}.npz"),
                            **{k: v.numpy() for k, v in cell_state(cell).items()})
        print(f"weights_{alias}.npz")


def make_rcnn_case():
    """RCNN.forward list semantics through the reference's own RCNN class (GS2D:128-190)."""
    mod = load_reference_module("gs2d")
    torch.manual_seed(5)
    low = torch.rand(1, 2, 6, 5)
    model = mod.RCNN(input_channels=2, hidden_channels=8, init_state_low=low, input_kernel_size=5,
                     step=7, effective_step=[0, 2, 3, 6])
    with torch.no_grad():
        for prm in model.crnn_cell.parameters():
            if prm.requires_grad and prm.dim() > 0:
                prm.mul_(30.0)  # Xavier*0.02 weights make the Pi term invisible; scale it up
        outputs, second_last = model()
    rec = {"init_state_low": low.numpy(), "step": np.array(7), "effective_step": np.array([0, 2, 3, 6]),
           "outputs": torch.cat(outputs, 0).numpy(), "second_last": second_last.numpy()}
    for k, v in model.state_dict().items():
        rec["state/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "rcnn_gs2d.npz"), **rec)
    print("rcnn_gs2d.npz", rec["outputs"].shape)


if __name__ == "__main__":
    torch.set_num_threads(1)  # oneDNN summation order depends on the thread count
    if "--data-loss-only" in sys.argv:
        make_data_loss_cases()
        sys.exit(0)
    if "--phys-only" in sys.argv:
        make_phys_cases()
        sys.exit(0)
    if "--rk4-only" in sys.argv:
        make_rk4_cases()
        sys.exit(0)
    make_case("fwd", (20, 24), 6, 11, amp=(-0.8, 0.8))
    make_case("gs2d", (20, 24), 6, 12)
    make_case("gs3d", (6, 8, 12), 5, 13)
    make_case("gs3d", (6, 16, 128), 3, 14, tag="gs3d_tma")
    make_case("bur1", (16, 20), 4, 15, amp=(-0.5, 0.5))
    make_case("lo1", (16, 20), 4, 16, amp=(-0.8, 0.8))
    make_case("bur3", (20, 24), 6, 17, amp=(-0.5, 0.5))
    make_case("lo3", (20, 24), 6, 18, amp=(-0.8, 0.8))
    make_case("lo3n", (20, 24), 6, 19, amp=(-0.8, 0.8))
    make_weights()
    # (rcnn_*.npz, train_*.npz and the Stage-3 checkpoint fixture come from make_golden_modules.py)
    make_data_loss_cases()
    make_phys_cases()
    make_rk4_cases()
