"""Shared pieces of the Stage-3 (coefficient fine-tuning) scripts: physics-based cells in fp64."""
import torch
import torch.nn as nn

from .. import _lib
from ..cells import Conv2dDerivative, FusedRCNN, PhysicsCell, derivative_table, laplace_table
from ..engine import CellSpec
from ._stage1 import upscaler  # noqa: F401  (BUR3:38-52 is the same upscaler)


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
        self.UpconvBlock = upscaler()
        self._setup(self.cell_cls(input_channels=input_channels, hidden_channels=hidden_channels,
                                  output_channels=output_channels, input_kernel_size=input_kernel_size,
                                  input_stride=input_stride, input_padding=input_padding), step, effective_step)


def _scalar(v):
    return nn.Parameter(torch.tensor(v, dtype=torch.float64), requires_grad=True)
