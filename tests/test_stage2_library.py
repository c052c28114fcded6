"""Stage-2 library of candidate terms: oracle and fused kernels against vectors recorded from the
reference's own `Loss_generator` (Stage-2/derivatives.py) and PDE_FIND_u.py's column construction."""
import os

import numpy as np
import pytest
import torch

from oracle import percnn_oracle as po
from tests.helpers import GOLDEN

NAMES = ("f_u", "f_v", "u", "v", "u_t", "v_t", "u_x", "u_y", "v_x", "v_y", "lap_u", "lap_v")


def _rel(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("kind", ["burgers", "lo"])
def test_oracle_library_matches_the_reference(kind):
    z = np.load(os.path.join(GOLDEN, f"stage2_{kind}.npz"))
    lib = po.stage2_library_torch(torch.from_numpy(z["output"]), kind, float(z["dt"]), float(z["dx"]))
    for n in NAMES + ("ones",):
        assert lib[n].shape == z["term/" + n].shape
        assert _rel(lib[n].numpy(), z["term/" + n]) <= 1e-6, n
    lhs, rhs = po.stage2_theta_np({k: torch.from_numpy(z["term/" + k]) for k in NAMES + ("ones",)}, z["idx"])
    assert [a + "*" + b for a in po.STAGE2_LIST_A for b in po.STAGE2_LIST_B] == list(z["lib"])
    assert _rel(lhs, z["lhs"]) <= 1e-14 and np.array_equal(rhs[:, 0:1], z["rhs_u"]) and np.array_equal(rhs[:, 1:2], z["rhs_v"])


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["burgers", "lo"])
def test_fused_library_matches_the_reference(kind):
    from percnn_b200 import library
    z = np.load(os.path.join(GOLDEN, f"stage2_{kind}.npz"))
    out = torch.from_numpy(z["output"]).cuda()
    lg = library.Loss_generator(dt=float(z["dt"]), dx=float(z["dx"]), kind=kind)
    assert library.gen_library() == list(z["lib"])
    pad = torch.cat((out[:, :, :, -2:], out, out[:, :, :, 0:3]), dim=3)
    pad = torch.cat((pad[:, :, -2:, :], pad, pad[:, :, 0:3, :]), dim=2)
    lib = (lg.get_phy_residual if kind == "burgers" else lg.get_library)(pad)
    for n in NAMES + ("ones",):
        assert tuple(lib[n].shape) == z["term/" + n].shape
        # fp32 stencils of O(1) data divided by dx (1e-2) or dx^2: compare relative to the term's own scale
        assert _rel(lib[n].cpu().numpy(), z["term/" + n]) <= 2e-5, n
    mu, mv = lg.get_residual_mse(out)
    assert abs(mu.item() - float(z["mse_u"])) <= 1e-4 * float(z["mse_u"]) and abs(mv.item() - float(z["mse_v"])) <= 1e-4 * float(z["mse_v"])
    # the 70-column matrix from the REFERENCE's terms (so the comparison isolates the column products): exact to fp64 rounding
    terms = torch.stack([torch.from_numpy(z["term/" + n]) for n in NAMES]).cuda()
    lhs, rhs = lg.theta(terms, torch.from_numpy(z["idx"]))
    assert _rel(lhs.cpu().numpy(), z["lhs"]) <= 1e-14
    assert np.array_equal(rhs[:, 0:1].cpu().numpy(), z["rhs_u"]) and np.array_equal(rhs[:, 1:2].cpu().numpy(), z["rhs_v"])
    with pytest.raises(ValueError):
        lg.get_phy_residual(torch.zeros_like(pad).normal_())     # not a periodic padding


@pytest.mark.gpu
def test_fused_library_fp64_matches_the_oracle_on_a_ragged_grid():
    from percnn_b200 import library
    g = torch.Generator().manual_seed(5)
    out = torch.rand((5, 2, 9, 14), generator=g, dtype=torch.float64)
    want = po.stage2_library_torch(out, "lo", 0.0125, 0.2)
    lg = library.Loss_generator(dt=0.0125, dx=0.2, kind="lo")
    got = lg.library_from_periodic(out.cuda())
    for n in NAMES:
        assert _rel(got[n].cpu().numpy(), want[n].numpy()) <= 1e-7, n   # the oracle keeps the reference's fp32 tap table
