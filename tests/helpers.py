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
