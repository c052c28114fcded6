"""Multi-GPU tests (need >= 2 CUDA devices, skipped otherwise): the slab-decomposed rollout must reproduce the
single-GPU rollout bit for bit (forward, every transport) and its gradients (taped forward + fused-halo adjoint
+ one all-reduce of the parameter sums).  Each case is a torchrun of scripts/check_slab*.py on 127.0.0.1."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, args, port, nproc=2, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", script)] + args
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    return res.returncode, res.stdout + res.stderr


needs2 = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")


@needs2
@pytest.mark.parametrize("transport", ["fused", "symm", "nccl"])
def test_slab_forward_is_bitwise_equal_to_single_gpu(transport):
    rc, out = _torchrun("check_slab.py", ["--shape", "40", "48", "256", "--steps", "9", "--transport", transport], 29611)
    assert "bitwise_equal=True" in out, out[-2000:]
    assert rc == 0


@needs2
def test_fused_halo_stress_full_width_planes():
    """Regression for the stage-release race: with 512^2 planes (all 148 CTAs busy, peer stores back-pressuring
    the LSU) an mbarrier arrive used to overtake a queued shared-memory load; ~1 rollout in 40 was off by 1e-5."""
    rc, out = _torchrun("check_slab.py", ["--shape", "128", "512", "512", "--steps", "2", "--repeat", "60",
                                          "--transport", "fused"], 29613, timeout=600)
    assert "bitwise_equal=True" in out, out[-2000:]


@needs2
@pytest.mark.parametrize("shape,steps", [(("32", "128", "128"), 23), (("16", "16", "128"), 12), (("128", "128", "128"), 50)])
def test_time_blocked_small_slab_rollout_over_nvlink(shape, steps):
    """The communication-avoiding persistent kernel (2K ghost planes every K steps, landing buffers alternating between
    two pairs) between real peers, repeated so that rank skew gets its chance: 16 / 8 / 64 planes per rank."""
    rc, out = _torchrun("check_slab.py", ["--shape", *shape, "--steps", str(steps), "--repeat", "25", "--transport", "fused"], 29614)
    assert "bitwise_equal=True" in out, out[-2000:]
    assert rc == 0


@needs2
def test_slab_training_step_matches_single_gpu_autograd():
    rc, out = _torchrun("check_slab_bwd.py", ["--shape", "64", "48", "128", "--steps", "6"], 29612)
    assert "ok=True" in out, out[-2000:]
