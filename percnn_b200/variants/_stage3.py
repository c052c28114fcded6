"""Shared pieces of the Stage-3 (coefficient fine-tuning) scripts: physics-based cells in fp64."""
import torch
import torch.nn as nn

from .. import _lib
from ..cells import Conv2dDerivative, FusedRCNN, PhysicsCell, derivative_table, laplace_table
from ..engine import CellSpec
from ._stage1 import get_ic_loss, upscaler  # noqa: F401  (BUR3:38-52, 487-496: the same upscaler and IC loss)


class Stage3RCNN(FusedRCNN):
    cell_cls = None

    def __init__(self, input_channels, hidden_channels, output_channels, init_state_low, input_kernel_size,
                 input_stride, input_padding, step=1, effective_step=[1]):
        super().__init__()
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.output_channels = output_channels
        self.input_kernel_size = input_kernel_size
        self.input_stride = input_stride
        self.input_padding = input_padding
        self.init_state_low = init_state_low
        self.init_state = []
        self.UpconvBlock = upscaler().double()   # the Stage-3 scripts run with torch.set_default_dtype(float64) (BUR3:15)
        self._setup(self.cell_cls(input_channels=input_channels, hidden_channels=hidden_channels,
                                  output_channels=output_channels, input_kernel_size=input_kernel_size,
                                  input_stride=input_stride, input_padding=input_padding), step, effective_step)


    # Checkpoints written by older versions of the scripts carry coefficients the current cell no longer has (the
    # shipped Burgers Stage-3 checkpoint: crnn_cell.C3_u, C4_u, C3_v, C4_v -- SURVEY 8c).  They are dropped, with a
    # warning, instead of failing the strict load; every key the model DOES have must still be present.
    STALE_KEY_PATTERN = r"^crnn_cell\.C[0-9]+_[uv]$"

    def load_state_dict(self, state_dict, strict=True, **kw):
        import re
        import warnings
        own = set(self.state_dict().keys())
        stale = [k for k in state_dict if k not in own and re.match(self.STALE_KEY_PATTERN, k)]
        if stale:
            warnings.warn(f"ignoring stale checkpoint coefficients the current cell does not have: {sorted(stale)}")
            state_dict = {k: v for k, v in state_dict.items() if k not in stale}
        return super().load_state_dict(state_dict, strict=strict, **kw)


def _scalar(v):
    return nn.Parameter(torch.tensor(v, dtype=torch.float64), requires_grad=True)
