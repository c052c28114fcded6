"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

#: golden file tag -> oracle variant name
GOLDEN_CASES = {
    "fwd": "fwd", "gs2d": "gs2d", "gs3d": "gs3d", "gs3d_tma": "gs3d", "bur1": "bur1", "lo1": "lo1",
    "bur3": "bur3", "lo3": "lo3", "lo3n": "lo3",
}


def load_golden(tag):
    z = np.load(os.path.join(GOLDEN, f"cell_{tag}.npz"))
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    grads = {k[len("grad/"):]: z[k] for k in z.files if k.startswith("grad/")}
    return z, params, grads


#: data-loss golden tag -> oracle variant name (tests/golden/dloss_*.npz, make_golden.make_data_loss_case)
DLOSS_CASES = {"gs2d": "gs2d", "gs3d": "gs3d", "gs3d_tma": "gs3d", "bur1": "bur1", "bur3": "bur3", "lo3": "lo3"}


def load_dloss(tag):
    z = np.load(os.path.join(GOLDEN, f"dloss_{tag}.npz"))
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    grads = {k[len("grad/"):]: z[k] for k in z.files if k.startswith("grad/")}
    return z, params, grads


def load_weights(alias):
    z = np.load(os.path.join(GOLDEN, f"weights_{alias}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def rel_linf(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def make_cell(tag):
    """Construct the percnn_b200 drop-in cell that corresponds to a golden tag / oracle variant."""
    from percnn_b200.variants import (burgers_stage1, burgers_stage3, gs2d, gs3d, lambda_omega_fwd, lo_stage1,
                                      lo_stage3)
    if tag == "fwd":
        return lambda_omega_fwd.RCNNCell(input_kernel_size=1, input_stride=1, input_padding=0)
    if tag == "gs2d":
        return gs2d.RCNNCell(input_channels=2, hidden_channels=8, input_kernel_size=5)
    if tag in ("gs3d", "gs3d_tma"):
        return gs3d.RCNNCell(input_channels=2, hidden_channels=2, input_kernel_size=5)
    kw = dict(input_channels=2, hidden_channels=4, output_channels=2, input_kernel_size=5, input_stride=1,
              input_padding=2)
    if tag == "bur1":
        return burgers_stage1.RCNNCell(**kw)
    if tag == "lo1":
        return lo_stage1.RCNNCell(**kw)
    if tag == "bur3":
        return burgers_stage3.RCNNCell(**kw)
    if tag == "lo3":
        return lo_stage3.RCNNCell(**kw)
    if tag == "lo3n":
        return lo_stage3.RCNNCellNoisy(**kw)
    raise KeyError(tag)


def cell_params_dict(cell):
    """{state_dict key: tensor on CPU} -- the oracle's parameter format."""
    return {k: v.detach().cpu() for k, v in cell.state_dict().items()}


def state_checksum(t):
    """Fingerprint of a seeded synthetic state (sum, sum |.|, sum of squares and 9 equally spaced samples): the full-size
    goldens store it instead of the multi-MB state itself, the tests regenerate the state from its seed and compare."""
    d = t.detach().double().reshape(-1)
    idx = torch.linspace(0, d.numel() - 1, 9, dtype=torch.float64).long()
    return np.concatenate([[float(d.sum()), float(d.abs().sum()), float((d * d).sum())], d[idx].numpy()])


def checksums_match(a, b):
    """The three sums depend on the summation order (thread count / device), so they are compared relative to the
    sum of magnitudes; the nine samples must agree to rounding."""
    a, b = np.asarray(a), np.asarray(b)
    scale = max(abs(float(b[1])), 1e-300)
    return bool(np.all(np.abs(a[:3] - b[:3]) <= 1e-11 * max(scale, abs(float(b[2])))) and np.allclose(a[3:], b[3:], rtol=1e-12, atol=1e-14))
