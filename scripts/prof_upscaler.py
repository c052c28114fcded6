"""A few forward + backward passes of the fused GS3D upscaler at the reference size (for an ncu launch list)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from percnn_b200.variants import gs2d, gs3d  # noqa: E402

for mod, low in ((gs3d, (24, 24, 24)), (gs2d, (25, 25))):
    m = mod.upscaler().cuda()
    x = torch.rand((1, 2, *low), device="cuda")
    with torch.no_grad():
        g = torch.rand_like(m(x))
    for _ in range(3):
        m.zero_grad(set_to_none=True)
        (m(x) * g).sum().backward()
    torch.cuda.synchronize()
