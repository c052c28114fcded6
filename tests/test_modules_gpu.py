"""Module-level parity: the drop-in `RCNN` classes of every script against the reference's OWN classes, and the
scripts' own `train()` loops reproduced through the drop-ins (tests/golden/make_golden_modules.py recorded the
reference side in the build container)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch import nn, optim
from torch.optim.lr_scheduler import StepLR

from percnn_b200.variants import (burgers_stage1, burgers_stage3, gs2d, gs3d, lambda_omega_fwd, lo_stage1, lo_stage3)
from tests.helpers import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = False      # the stock upscaler convs must not silently run in TF32
torch.backends.cuda.matmul.allow_tf32 = False

STAGE_KW = dict(input_channels=2, hidden_channels=4, output_channels=2, input_kernel_size=5, input_stride=1, input_padding=2)


def _build(alias, low, step, eff):
    if alias == "fwd":
        return lambda_omega_fwd.RCNN(input_kernel_size=1, ini_state=low.cpu().numpy(), input_stride=1, input_padding=0, step=step,
                                     effective_step=eff)
    if alias == "gs2d":
        return gs2d.RCNN(input_channels=2, hidden_channels=8, init_state_low=low, input_kernel_size=5, step=step, effective_step=eff)
    if alias == "gs3d":
        return gs3d.RCNN(input_channels=2, hidden_channels=2, init_state_low=low, input_kernel_size=5, step=step, effective_step=eff)
    mod = {"bur1": burgers_stage1, "lo1": lo_stage1, "bur3": burgers_stage3, "lo3": lo_stage3}[alias]
    return mod.RCNN(init_state_low=low, step=step, effective_step=eff, **STAGE_KW)


@pytest.mark.parametrize("alias", ["fwd", "gs2d", "gs3d", "bur1", "lo1", "bur3", "lo3"])
def test_rcnn_forward_and_gradients_match_the_reference_class(alias):
    z = np.load(os.path.join(GOLDEN, f"rcnn_{alias}.npz"))
    low = torch.from_numpy(z["init_state_low"]).to(DEV)
    step, eff = int(z["step"]), [int(s) for s in z["effective_step"]]
    model = _build(alias, low, step, eff).to(DEV)
    if alias == "fwd":
        model.init_state = model.init_state.to(DEV)
    sd = {k[len("state/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state/")}
    model.load_state_dict(sd, strict=True)
    f64 = z["outputs"].dtype == np.float64
    tol, gtol = (1e-11, 1e-9) if f64 else (1e-5, 3e-5)
    with torch.no_grad():
        outputs, second_last = model()
    assert isinstance(outputs, list) and len(outputs) == 1 + len(eff)
    assert rel_l2(torch.cat(tuple(outputs), 0).cpu().numpy(), z["outputs"]) <= tol
    assert rel_l2(second_last.cpu().numpy(), z["second_last"]) <= tol
    outputs, _ = model()
    out = torch.cat(tuple(outputs), dim=0)
    loss = out[1:].pow(2).mean()
    loss.backward(retain_graph=True)
    assert abs(loss.item() - float(z["loss"])) <= (1e-11 if f64 else 1e-5) * abs(float(z["loss"]))
    named = dict(model.named_parameters())
    checked = 0
    for key in [k for k in z.files if k.startswith("grad/")]:
        name = key[len("grad/"):]
        got = named[name].grad
        assert got is not None, name
        ref = z[key]
        scalar_fp32 = ref.ndim == 0 and not f64
        err = rel_l2(got.cpu().numpy(), ref)
        assert err <= (3e-4 if scalar_fp32 else gtol), (name, err)
        checked += 1
    assert checked >= 8


def _get_ic_loss_gs2d(model):
    """GS2D:331-338 through the drop-in of the same name (fused upscaler + fused MSE); the stock formulation on the same
    model must agree (it differs only in who computes the mean)."""
    loss = gs2d.get_ic_loss(model)
    with torch.no_grad():
        init_state_bicubic = F.interpolate(model.init_state_low, (100, 100), mode="bicubic")
        stock = nn.MSELoss()(model.UpconvBlock(model.init_state_low), init_state_bicubic)
    assert abs(loss.item() - stock.item()) <= 2e-6 * abs(stock.item())
    return loss


def test_train_loop_of_train_2drd_reproduces_the_reference_loss_trajectory(tmp_path):
    """GS2D:374-425 through the drop-in classes: Adam + StepLR, 40 * data + 0.25 * ic, backward(retain_graph=True);
    then the checkpoint round trip of GS2D:414-422 / 432-439."""
    z = np.load(os.path.join(GOLDEN, "train_gs2d.npz"))
    low = torch.from_numpy(z["init_state_low"]).to(DEV)
    step = int(z["step"])
    gt = torch.from_numpy(z["truth_sub"]).to(DEV)

    def fresh():
        m = gs2d.RCNN(input_channels=2, hidden_channels=8, init_state_low=low, input_kernel_size=5, step=step,
                      effective_step=list(range(step))).to(DEV)
        return m

    model = fresh()
    model.load_state_dict({k[len("init/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("init/")}, strict=True)
    optimizer = optim.Adam(model.parameters(), lr=float(z["lr"]))
    scheduler = StepLR(optimizer, step_size=200, gamma=0.985)
    loss_func = gs2d.loss_generator(0.5, 0.01)

    def one_epoch(model, optimizer, scheduler):
        optimizer.zero_grad()
        output, _ = model()
        output = torch.cat(tuple(output), dim=0)
        mse_loss = nn.MSELoss()
        pred = output[0:-1:20, :, ::4, ::4]
        idx = int(pred.shape[0] * 0.9)
        loss_data = mse_loss(pred[:idx], gt[:idx])
        loss_valid = mse_loss(pred[idx:], gt[idx:])
        loss_ic = _get_ic_loss_gs2d(model)
        loss_phy = gs2d.loss_gen(output, loss_func)
        loss = 40 * loss_data + 0.25 * loss_ic
        loss.backward(retain_graph=True)
        optimizer.step()
        scheduler.step()
        return loss.item(), (loss_ic.item(), loss_data.item(), loss_valid.item(), loss_phy.item())

    losses, printed = [], []
    for _ in range(len(z["losses"])):
        l, pr = one_epoch(model, optimizer, scheduler)
        losses.append(l)
        printed.append(pr)
    assert np.allclose(losses[0], z["losses"][0], rtol=2e-5)
    assert np.allclose(losses, z["losses"], rtol=5e-4), (losses, z["losses"])
    assert np.allclose(np.array(printed), z["printed"], rtol=2e-3, atol=1e-7), (printed, z["printed"])
    for k in [k for k in z.files if k.startswith("final/")]:
        got = model.state_dict()[k[len("final/"):]].cpu().numpy()
        assert np.allclose(got, z[k], rtol=2e-3, atol=2e-5), k
    # checkpoint round trip: save model + optimizer, load into a fresh model, continue identically
    path = os.path.join(tmp_path, "checkpoint.pt")
    torch.save({"model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict()}, path)
    ck = torch.load(path)
    model2 = fresh()
    model2.load_state_dict(ck["model_state_dict"])
    opt2 = optim.Adam(model2.parameters(), lr=0.0)
    opt2.load_state_dict(ck["optimizer_state_dict"])
    sch2 = StepLR(opt2, step_size=200, gamma=0.98)
    l1, _ = one_epoch(model, optimizer, StepLR(optimizer, step_size=200, gamma=0.98))
    l2, _ = one_epoch(model2, opt2, sch2)
    assert abs(l1 - l2) <= 1e-6 * abs(l1)      # (torch's transposed-conv backward is not bit-reproducible run to run)


def test_train_loop_of_percnn_LO_eqn_reproduces_the_reference_loss_trajectory():
    """FWD:360-383: physics loss only (fused), fp64, `model.init_state` re-assigned every epoch."""
    z = np.load(os.path.join(GOLDEN, "train_fwd.npz"))
    ini = z["ini_state"]
    step = int(z["step"])
    model = lambda_omega_fwd.RCNN(input_kernel_size=1, ini_state=ini, input_stride=1, input_padding=0, step=step,
                                  effective_step=list(range(step))).to(DEV)
    model.load_state_dict({k[len("init/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("init/")}, strict=True)
    optimizer = optim.Adam(model.parameters(), lr=float(z["lr"]))
    scheduler = StepLR(optimizer, step_size=25, gamma=0.98)
    loss_func = lambda_omega_fwd.loss_generator(0.0125, 0.2)
    losses = []
    for _ in range(len(z["losses"])):
        optimizer.zero_grad()
        model.init_state = torch.tensor(ini, dtype=torch.float64).cuda()
        output, _ = model()
        output = torch.cat(tuple(output), dim=0)
        loss = lambda_omega_fwd.loss_gen(output, loss_func)
        loss.backward(retain_graph=True)
        optimizer.step()
        scheduler.step()
        losses.append(loss.item())
    assert np.allclose(losses, z["losses"], rtol=1e-8), (losses, z["losses"])
    for k in [k for k in z.files if k.startswith("final/")]:
        got = model.state_dict()[k[len("final/"):]].cpu().numpy()
        assert np.allclose(got, z[k], rtol=1e-7, atol=1e-10), k


@pytest.mark.parametrize("alias,mod", [("fwd", lambda_omega_fwd), ("gs2d", gs2d), ("gs3d", gs3d)])
def test_get_phy_loss_returns_the_reference_residual_fields(alias, mod):
    """`loss_generator.get_phy_Loss` on the periodically padded trajectory (GS2D:270-329): mse(f_u) + mse(f_v) must be
    the reference's `loss_gen` value recorded in tests/golden/phys_*.npz."""
    z = np.load(os.path.join(GOLDEN, f"phys_{alias}.npz"))
    traj = torch.from_numpy(z["traj"]).to(DEV)
    out = traj
    for ax in range(2, traj.dim()):
        n = out.shape[ax]
        out = torch.cat((out.narrow(ax, n - 2, 2), out, out.narrow(ax, 0, 3)), dim=ax)
    gen = mod.loss_generator()
    f_u, f_v = gen.get_phy_Loss(out)
    assert tuple(f_u.shape) == (traj.shape[0] - 2, 1, *[n + 1 for n in traj.shape[2:]])
    loss = (f_u ** 2).mean() + (f_v ** 2).mean()
    tol = 1e-10 if traj.dtype == torch.float64 else 2e-5
    assert abs(loss.item() - float(z["loss"])) <= tol * abs(float(z["loss"]))
    with pytest.raises(ValueError):
        gen.get_phy_Loss(out + torch.rand_like(out))
