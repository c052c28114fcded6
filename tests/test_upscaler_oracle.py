"""The oracle's upscaler / IC-loss restatements against vectors recorded from the reference's own `upscaler` classes and
`get_ic_loss` functions (tests/golden/make_golden_upscaler.py).  CPU."""
import os

import numpy as np
import pytest
import torch

from oracle import percnn_oracle as po
from tests.helpers import GOLDEN, rel_l2

KIND = {"gs2d": "gs2d", "gs3d": "gs3d", "bur1": "stage", "bur3": "stage"}


def _load(alias):
    z = np.load(os.path.join(GOLDEN, f"up_{alias}.npz"))
    sd = {k[len("state/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state/")}
    return z, sd


@pytest.mark.parametrize("alias", ["gs2d", "gs3d", "bur1", "bur3"])
def test_upscaler_restatements_match_the_reference_module(alias):
    z, sd = _load(alias)
    f64 = z["out"].dtype == np.float64
    low = torch.from_numpy(z["low"])
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = po.upscaler_torch(low, params, KIND[alias])
    assert rel_l2(out.detach().numpy(), z["out"]) <= (1e-13 if f64 else 1e-6)
    assert rel_l2(po.upscaler_np(low, sd, KIND[alias]), z["out"]) <= (1e-13 if f64 else 2e-6)
    (out * torch.from_numpy(z["gout"])).sum().backward()
    for k in [k for k in z.files if k.startswith("grad/")]:
        name = k[len("grad/"):]
        assert rel_l2(params[name].grad.numpy(), z[k]) <= (1e-12 if f64 else 2e-5), name


@pytest.mark.parametrize("alias", ["gs2d", "gs3d", "bur1", "bur3"])
def test_ic_loss_restatement_matches_get_ic_loss(alias):
    z, sd = _load(alias)
    f64 = z["out"].dtype == np.float64
    low = torch.from_numpy(z["ic_low"])
    size = z["ic_target"].shape[2:]
    tgt = po.ic_target_torch(low, KIND[alias], size)
    assert rel_l2(tgt.numpy(), z["ic_target"]) <= (1e-13 if f64 else 1e-6)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss = po.ic_loss_torch(low, params, KIND[alias], size)
    assert abs(loss.item() - float(z["ic_loss"])) <= (1e-12 if f64 else 1e-5) * abs(float(z["ic_loss"]))
    loss.backward()
    for k in [k for k in z.files if k.startswith("ic_grad/")]:
        name = k[len("ic_grad/"):]
        assert rel_l2(params[name].grad.numpy(), z[k]) <= (1e-11 if f64 else 3e-5), name
