"""Golden vectors for the Stage-2 library from the reference's OWN `Loss_generator` (derivatives.py) and the column
construction of PDE_FIND_u.py (build container only).

    python tests/golden/make_golden_stage2.py          # rewrites tests/golden/stage2_*.npz

The reference modules are imported from /root/reference (stubs for matplotlib / scipy.io file loading are not needed:
only classes and functions are used, `__main__` never runs); `.cuda()` is a no-op here.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

DIRS = {"burgers": "DataDrivenDiscoveryOfPDEs/2D_Burgers_eqn/Stage-2", "lo": "DataDrivenDiscoveryOfPDEs/2D_Lambda_Omega_eqn/stage-2"}
CONSTS = {"burgers": (0.00025, 1.0 / 100), "lo": (0.0125, 0.2)}


def load(dirname, name):
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    path = os.path.join(mg.REF, dirname)
    sys.path.insert(0, path)
    try:
        spec = importlib.util.spec_from_file_location(f"percnn_ref_s2_{name}_{abs(hash(dirname)) % 1000}", os.path.join(path, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(path)
        sys.modules.pop("derivatives", None)
        torch.set_default_dtype(torch.float32)
    return mod


def make_case(kind, shape, seed):
    der = load(DIRS[kind], "derivatives")
    pde = load(DIRS[kind], "PDE_FIND_u")
    dt, dx = CONSTS[kind]
    T, H, W = shape
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H) * (2 * np.pi / H), torch.arange(W) * (2 * np.pi / W), indexing="ij")
    frames = []
    for t in range(T):       # smooth periodic fields drifting in time + a little noise (so every term is non-trivial)
        ph = 0.05 * t
        u = 0.6 * torch.sin(xx + ph) * torch.cos(2 * yy - ph) + 0.2 * torch.cos(3 * xx + yy)
        v = 0.5 * torch.cos(xx - 2 * ph) * torch.sin(yy + ph) - 0.3 * torch.sin(2 * xx - yy)
        frames.append(torch.stack((u, v)))
    output = (torch.stack(frames) + 0.01 * torch.randn((T, 2, H, W), generator=g)).float()
    lg = der.Loss_generator(dt=dt, dx=dx)
    mse_u, mse_v = lg.get_residual_mse(output)
    pad = torch.cat((output[:, :, :, -2:], output, output[:, :, :, 0:3]), dim=3)
    pad = torch.cat((pad[:, :, -2:, :], pad, pad[:, :, 0:3, :]), dim=2)
    terms = (lg.get_phy_residual if kind == "burgers" else lg.get_library)(pad)
    rec = {"output": output.numpy(), "dt": np.array(dt), "dx": np.array(dx), "mse_u": np.array(mse_u.item()), "mse_v": np.array(mse_v.item())}
    for k, v in terms.items():
        rec["term/" + k] = v.detach().numpy()
    # PDE_FIND_u.py:228-259, verbatim in effect: to_numpy_float64, sampled rows, eval of every library expression
    terms_dict = der.to_numpy_float64(dict(terms))
    lib = pde.gen_library()
    n = terms_dict["u"].shape[0]
    idx = np.random.RandomState(seed).choice(n, int(n * 0.2), replace=False)
    scope = {k: v[idx, :] for k, v in terms_dict.items()}
    lhs = np.concatenate([eval(e, {}, scope) for e in lib], axis=1)
    rec.update({"lib": np.array(lib), "idx": idx, "lhs": lhs, "rhs_u": scope["u_t"], "rhs_v": scope["v_t"]})
    np.savez_compressed(os.path.join(HERE, f"stage2_{kind}.npz"), **rec)
    print(f"stage2_{kind}.npz terms {tuple(terms['u'].shape)} lhs {lhs.shape} mse {mse_u.item():.4g} {mse_v.item():.4g}")


if __name__ == "__main__":
    make_case("burgers", (7, 20, 24), 71)
    make_case("lo", (6, 17, 13), 72)
