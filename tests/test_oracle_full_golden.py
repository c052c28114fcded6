"""CPU: the oracle (oracle/percnn_oracle.py) against the FULL-SIZE vectors recorded from the reference's own classes
(tests/golden/full_*.npz, made by tests/golden/make_golden_full.py): pins the restatement at BASELINE.json's config
sizes and rollout lengths, not only on the small round-1 grids.  Sized to stay within ~1 minute of CPU."""
import os

import numpy as np
import torch

from oracle import percnn_oracle as po
from tests.helpers import GOLDEN, checksums_match, load_weights, rel_l2, rel_linf, state_checksum


def _full(name):
    return np.load(os.path.join(GOLDEN, f"full_{name}.npz"))


def _roll(h0, params, variant, nsteps, keep):
    out = {}
    h = h0
    with torch.no_grad():
        for s in range(nsteps):
            h = po.cell_step_torch(h, params, variant)
            if s + 1 in keep:
                out[s + 1] = h.clone()
    return out


def test_seeded_initial_states_reproduce_the_golden_checksums():
    assert checksums_match(state_checksum(po.ic_spiral_2d(128)), _full("cfg1")["h0_checksum"])
    assert checksums_match(state_checksum(po.ic_gs_2d(256, seed=0)), _full("cfg2")["h0_checksum"])
    assert checksums_match(state_checksum(po.ic_fourier_2d(512, seed=1)), _full("cfg3_bur1")["h0_checksum"])
    assert checksums_match(state_checksum(po.ic_gs_3d((128, 128, 128), seed=0)), _full("cfg4")["h0_checksum"])


def test_oracle_cfg1_full_length():
    z = _full("cfg1")
    st = _roll(po.ic_spiral_2d(128), load_weights("fwd"), "fwd", 200, {1, 50, 200})
    for t in (1, 50, 200):
        assert rel_l2(st[t].numpy(), z[f"state_{t}"]) <= 1e-12, t
    # independent numpy restatement, same length
    h = po.ic_spiral_2d(128).numpy()
    p = load_weights("fwd")
    for _ in range(200):
        h = po.cell_step_np(h, p, "fwd")
    assert rel_l2(h, z["state_200"]) <= 1e-10


def test_oracle_cfg2_full_length():
    z = _full("cfg2")
    st = _roll(po.ic_gs_2d(256, seed=0), load_weights("gs2d"), "gs2d", 1000, {1, 100, 1000})
    assert rel_linf(st[1].numpy(), z["state_1"]) <= 1e-6
    assert rel_l2(st[100].numpy()[..., ::2, ::2], z["state_100"]) <= 2e-6
    # fp32 on a different thread count is not bit-reproducible (4.8e-7 at t = 1000)
    assert rel_l2(st[1000].numpy(), z["state_1000"]) <= 5e-6
    assert rel_l2(st[1000].numpy()[..., ::2, ::2], z["f64_state_1000"]) <= max(4 * float(z["ref32_vs_f64_l2"][2]), 2e-6)


def test_oracle_cfg3_burgers_advection_bptt():
    z = _full("cfg3_bur3")
    h0 = po.ic_fourier_2d(512, seed=1, dtype=torch.float64).requires_grad_(True)
    tgt = po.ic_fourier_2d(512, seed=2, dtype=torch.float64)[:, :, ::2, ::2].expand(8, -1, -1, -1)
    p = {k: v.clone().requires_grad_(v.dim() == 0) for k, v in po.make_phys_params("bur3").items()}
    outs, _ = po.rollout_torch(h0, p, "bur3", 40, range(40))
    out = torch.cat(outs, 0)
    loss = torch.mean((out[0:-1:5, :, ::2, ::2] - tgt) ** 2)
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) <= 1e-12 * abs(float(z["loss"]))
    assert rel_l2(h0.grad.numpy()[..., ::4, ::4], z["g_h0_sub"]) <= 1e-11
    for key in [k for k in z.files if k.startswith("grad/")]:
        assert rel_l2(p[key[len("grad/"):]].grad.numpy(), z[key]) <= 1e-10, key


def test_oracle_cfg4_first_50_steps():
    z = _full("cfg4")
    st = _roll(po.ic_gs_3d((128, 128, 128), seed=0), load_weights("gs3d"), "gs3d", 50, {1, 50})
    for t in (1, 50):
        s = st[t].numpy()
        assert rel_l2(s[..., ::4, ::4, ::4], z[f"state_{t}_sub"]) <= 2e-6, t
        assert rel_l2(s[:, :, 64], z[f"state_{t}_plane64"]) <= 2e-6, t
