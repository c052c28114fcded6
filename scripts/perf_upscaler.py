"""Timing of the fused initial-state generator (forward, forward + adjoint) next to the stock cuDNN modules."""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from percnn_b200.variants import burgers_stage1, gs2d, gs3d  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
DEV = "cuda:0"


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def stock(m):
    return m.convnet


for name, mod, low in (("GS2D 25^2 -> 100^2", gs2d, (25, 25)), ("GS2D 128^2 -> 512^2", gs2d, (128, 128)),
                       ("BUR1 50^2 -> 100^2", burgers_stage1, (50, 50)), ("GS3D 24^3 -> 48^3", gs3d, (24, 24, 24)),
                       ("GS3D 64^3 -> 128^3", gs3d, (64, 64, 64)), ("GS3D 128^3 -> 256^3", gs3d, (128, 128, 128))):
    m = mod.upscaler().to(DEV)
    x = torch.rand((1, 2, *low), device=DEV)
    with torch.no_grad():
        g = torch.rand_like(m(x))

    def fwd(mm=m):
        with torch.no_grad():
            mm(x)

    def fb(mm=m):
        mm.zero_grad(set_to_none=True)
        (mm(x) * g).sum().backward()

    def sfwd():
        with torch.no_grad():
            stock(m)(x)

    def sfb():
        m.zero_grad(set_to_none=True)
        (stock(m)(x) * g).sum().backward()

    n = 5 if "256" in name else 20
    print(f"{name:22s} fused fwd {timeit(fwd, n):9.1f} us  f+b {timeit(fb, n):9.1f} us | stock cuDNN fwd {timeit(sfwd, n):9.1f} us  f+b {timeit(sfb, n):9.1f} us", flush=True)
