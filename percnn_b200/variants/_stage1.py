"""Shared pieces of the two Stage-1 (data reconstruction) scripts: 5x5 Pi convs with 16 channels."""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..cells import FusedRCNN, PiCell
from ..upscaler import FusedUpscaler, ic_loss


class upscaler(FusedUpscaler):
    """BUR1:38-52 / LO1:38-52: one stride-2 transposed conv, tanh, 1x1 conv.  The layers are registered both
    by name and inside `convnet`, which is why the shipped checkpoints carry aliased keys; they hold
    the parameters, the arithmetic runs in the fused kernels."""
    up_ndim, up_channels, up_act, up_stride2 = 2, 16, "tanh", 1

    def __init__(self):
        super().__init__()
        self.layers = []
        self.up0 = nn.ConvTranspose2d(2, 16, kernel_size=5, padding=2, stride=2, output_padding=1, bias=True)
        self.tanh = nn.Tanh()
        self.out = nn.Conv2d(16, 2, 1, 1, padding=0, bias=True)
        self.convnet = nn.Sequential(self.up0, self.tanh, self.out)

    def _up_modules(self):
        return [self.up0, self.out]


def get_ic_loss(model):
    """BUR1:462-471 (= LO1:450-459, BUR3:487-496): mse(UpconvBlock(init_state_low), bicubic align_corners interpolation
    of the periodically extended init_state_low, last row/column dropped), fused."""
    return ic_loss(model, "bicubic_periodic")


class Stage1Cell(PiCell):
    _dx = 0.01
    _dt = 0.00025
    _nu_up = 0.01

    def __init__(self, input_channels, hidden_channels, output_channels, input_kernel_size, input_stride, input_padding):
        super().__init__()
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels      # ignored by the reference too: 16 is hard-coded (BUR1:108)
        self.output_channels = output_channels
        self.input_kernel_size = 5
        self.input_stride = input_stride
        self.input_padding = input_padding
        self.nu_up = self._nu_up
        np.random.seed(1234)                         # BUR1:98
        ca, cb = np.random.rand(), np.random.rand()
        self._build(ndim=2, dtype=torch.float32, ksize=5, hidden=16, dx=self._dx, dt=self._dt,
                    coef_mode=_lib.COEF_SIGMOID, mu_up=self.nu_up, coef_names=("CA", "CB"), coef_init=(ca, cb),
                    init_scale=0.5, init_kind="uniform")


class Stage1RCNN(FusedRCNN):
    cell_cls = Stage1Cell

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
