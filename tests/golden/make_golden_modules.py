"""Module-level golden vectors from the reference's OWN `RCNN` classes and `train()` functions (build container only).

    python tests/golden/make_golden_modules.py        # rewrites tests/golden/rcnn_*.npz and train_*.npz

* rcnn_<alias>.npz  -- `RCNN.forward()` list semantics (outputs for the effective steps, second_last_state), the
  training-style loss on the concatenated outputs and its autograd gradients for EVERY parameter (cell + upscaler),
  for all seven scripts (FWD:124-218, GS2D:128-190, GS3D:151-214, BUR1:190-303, LO1:183-296, BUR3:243-356,
  LO3:240-353).
* train_gs2d.npz / train_fwd.npz -- the scripts' own `train()` (GS2D:374-425: Adam + StepLR, 40*data + 0.25*ic loss,
  `backward(retain_graph=True)`; FWD:360-383: physics loss only, fp64) run for a few iterations on a small synthetic
  problem: the loss of every iteration and the parameters afterwards.
"""
import contextlib
import io
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as mg  # noqa: E402

STAGE_KW = dict(input_channels=2, hidden_channels=4, output_channels=2, input_kernel_size=5, input_stride=1, input_padding=2)


def build_rcnn(alias, mod, low, step, eff):
    torch.set_default_dtype(mod._default_dtype)
    try:
        if alias == "fwd":
            return mod.RCNN(input_kernel_size=1, ini_state=low.numpy(), input_stride=1, input_padding=0, step=step, effective_step=eff)
        if alias == "gs2d":
            return mod.RCNN(input_channels=2, hidden_channels=8, init_state_low=low, input_kernel_size=5, step=step, effective_step=eff)
        if alias == "gs3d":
            return mod.RCNN(input_channels=2, hidden_channels=2, init_state_low=low, input_kernel_size=5, step=step, effective_step=eff)
        return mod.RCNN(init_state_low=low, step=step, effective_step=eff, **STAGE_KW)
    finally:
        torch.set_default_dtype(torch.float32)


def make_rcnn_case(alias, low_shape, seed):
    mod = mg.load_reference_module(alias)
    dtype = mod._default_dtype
    g = torch.Generator().manual_seed(seed)
    low = (torch.rand((1, 2, *low_shape), generator=g, dtype=torch.float64) * 0.8 + 0.1).to(dtype)
    step, eff = 7, [0, 2, 3, 6]
    torch.manual_seed(seed)
    model = build_rcnn(alias, mod, low, step, eff)
    cell = model.rcnn_cell if alias == "fwd" else model.crnn_cell
    with torch.no_grad():
        if alias in ("fwd", "gs2d", "gs3d", "bur1", "lo1"):
            for prm in cell.parameters():
                if prm.requires_grad and prm.dim() > 0:
                    prm.mul_(10.0 if alias in ("bur1", "lo1", "fwd") else 30.0)   # make the Pi term visible in a 7-step rollout
    outputs, second_last = model()
    out = torch.cat(tuple(outputs), dim=0)
    loss = out[1:].pow(2).mean()
    loss.backward()
    rec = {"init_state_low": low.numpy(), "step": np.array(step), "effective_step": np.array(eff),
           "outputs": out.detach().numpy(), "second_last": second_last.detach().numpy(), "loss": np.array(loss.item())}
    for k, v in model.state_dict().items():
        rec["state/" + k] = v.detach().numpy()
    seen = set()
    for n, prm in model.named_parameters():
        if prm.requires_grad and prm.grad is not None and id(prm) not in seen:
            seen.add(id(prm))
            rec["grad/" + n] = prm.grad.numpy()
    np.savez_compressed(os.path.join(HERE, f"rcnn_{alias}.npz"), **rec)
    print(f"rcnn_{alias}.npz outputs {tuple(out.shape)} loss {loss.item():.6g} grads {len(seen)}")


def make_train_gs2d():
    """GS2D:374-425 `train(model, truth, n_iters, time_batch_size, lr, dt, dx, cont=False)` for 4 iterations."""
    mod = mg.load_reference_module("gs2d")
    g = torch.Generator().manual_seed(3)
    low = torch.rand((1, 2, 25, 25), generator=g) * 0.6 + 0.2          # get_ic_loss interpolates to (100, 100) (GS2D:334)
    step = 41                                                           # 41 states + h0 -> output[0:-1:20] = frames 0, 20, 40
    truth = torch.rand((41, 2, 100, 100), generator=g) * 0.6 + 0.2
    torch.manual_seed(11)
    model = mod.RCNN(input_channels=2, hidden_channels=8, init_state_low=low, input_kernel_size=5, step=step,
                     effective_step=list(range(0, step)))
    init_sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        losses = mod.train(model, truth, 4, 41, 2e-3, 0.5, 0.01, cont=False)
    rec = {"init_state_low": low.numpy(), "truth_sub": truth[::20, :, ::4, ::4].numpy(), "step": np.array(step),
           "losses": np.array(losses), "lr": np.array(2e-3)}
    m = re.findall(r"ic_loss: ([0-9.eE+-]+), data_loss: ([0-9.eE+-]+), val_loss: ([0-9.eE+-]+), loss phy_loss: ([0-9.eE+-]+)", buf.getvalue())
    rec["printed"] = np.array([[float(x) for x in row] for row in m])
    for k, v in init_sd.items():
        rec["init/" + k] = v.numpy()
    for k, v in model.state_dict().items():
        rec["final/" + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "train_gs2d.npz"), **rec)
    print("train_gs2d.npz losses", losses, "printed", rec["printed"].shape)


def make_train_fwd():
    """FWD:360-383 `train(model, init_state, n_iters, lr, dt, dx, save_path)`: physics loss only, fp64, 5 iterations."""
    mod = mg.load_reference_module("fwd")
    torch.set_default_dtype(torch.float64)
    try:
        g = torch.Generator().manual_seed(4)
        ini = (torch.rand((1, 2, 24, 28), generator=g, dtype=torch.float64) - 0.5).numpy()
        torch.manual_seed(12)
        model = mod.RCNN(input_kernel_size=1, ini_state=ini, input_stride=1, input_padding=0, step=12, effective_step=list(range(12)))
        init_sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            mod.train(model, ini, 5, 2e-3, 0.0125, 0.2, "/tmp/")
    finally:
        torch.set_default_dtype(torch.float32)
    losses = [float(x) for x in re.findall(r"Epoch loss: ([0-9.eE+-]+)", buf.getvalue())]
    rec = {"ini_state": ini, "step": np.array(12), "losses": np.array(losses), "lr": np.array(2e-3)}
    for k, v in init_sd.items():
        rec["init/" + k] = v.numpy()
    for k, v in model.state_dict().items():
        rec["final/" + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "train_fwd.npz"), **rec)
    print("train_fwd.npz losses", losses)


def make_ckpt_bur3():
    """The shipped Stage-3 Burgers checkpoint's state_dict: it carries C3_*/C4_* keys of an older version of the script
    (SURVEY 8c), which the drop-in's load_state_dict must tolerate the way `strict=False` would."""
    ck = torch.load(os.path.join(mg.REF, "DataDrivenDiscoveryOfPDEs/2D_Burgers_eqn/Stage-3/model/checkpoint.pt"),
                    map_location="cpu", weights_only=False)
    np.savez_compressed(os.path.join(HERE, "ckpt_bur3_stage3.npz"), **{k: v.numpy() for k, v in ck["model_state_dict"].items()})


if __name__ == "__main__":
    torch.set_num_threads(1)
    make_ckpt_bur3()
    make_rcnn_case("fwd", (20, 24), 51)
    make_rcnn_case("gs2d", (6, 5), 5)
    make_rcnn_case("gs3d", (4, 5, 6), 52)
    make_rcnn_case("bur1", (10, 12), 53)
    make_rcnn_case("lo1", (10, 12), 54)
    make_rcnn_case("bur3", (10, 12), 55)
    make_rcnn_case("lo3", (10, 12), 56)
    make_train_gs2d()
    make_train_fwd()
