"""Drop-in for the classes of DataDrivenModeling/3d_gs_rd/train_3drd.py (GS3D:41-214)."""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib, losses
from ..cells import FusedRCNN, PiCell
from ..upscaler import FusedUpscaler, ic_loss


class upscaler(FusedUpscaler):
    """GS3D:41-56: stride-2 then stride-1 transposed conv, 1x1x1 conv; the registered layers hold the parameters
    (state_dict keys `convnet.{0,2,3}.*`), the arithmetic runs in the fused kernels."""
    up_ndim, up_channels, up_act, up_stride2 = 3, 8, "sigmoid", 1

    def __init__(self):
        super().__init__()
        self.layers = [
            nn.ConvTranspose3d(2, 8, kernel_size=5, padding=2, stride=2, output_padding=1, bias=True),
            nn.Sigmoid(),
            nn.ConvTranspose3d(8, 8, kernel_size=5, padding=2, stride=1, output_padding=0, bias=True),
            nn.Conv3d(8, 2, 1, 1, padding=0, bias=True),
        ]
        self.convnet = nn.Sequential(*self.layers)

    def _up_modules(self):
        return [self.convnet[0], self.convnet[2], self.convnet[3]]


class RCNNCell(PiCell):
    """GS3D:58-139: fp32, 13-point 3-D Laplacian, 1x1x1 Pi convs."""

    def __init__(self, input_channels, hidden_channels, input_kernel_size=5):
        super().__init__()
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.input_kernel_size = input_kernel_size
        self.input_stride = 1
        self.mu_up = 0.274
        np.random.seed(1234)             # GS3D:75
        ca, cb = (np.random.rand() - 0.5) * 2, (np.random.rand() - 0.5) * 2
        self._build(ndim=3, dtype=torch.float32, ksize=1, hidden=hidden_channels, dx=100 / 48, dt=0.5,
                    coef_mode=_lib.COEF_SIGMOID, mu_up=self.mu_up, coef_names=("CA", "CB"), coef_init=(ca, cb),
                    init_scale=0.01, init_kind="xavier")


class RCNN(FusedRCNN):
    def __init__(self, input_channels, hidden_channels, init_state_low, input_kernel_size, output_channels=1, step=1,
                 effective_step=None):
        super().__init__()
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.output_channels = 1
        self.input_kernel_size = input_kernel_size
        self.init_state_low = init_state_low
        self.init_state = []
        self.UpconvBlock = upscaler()
        self._setup(RCNNCell(input_channels=input_channels, hidden_channels=hidden_channels,
                             input_kernel_size=input_kernel_size), step, effective_step)


class loss_generator(losses.LossGenerator):
    """GS3D:264-327 `loss_generator(dt, dx)`: Gray-Scott residual with Du = 0.2, Dv = 0.1, f = 0.025, k = 0.055."""

    def __init__(self, dt=0.5, dx=(100 / 48)):
        super().__init__(losses.gray_scott_spec(0.2, 0.1, 0.025, 0.055, dt, dx))


def get_ic_loss(model):
    """GS3D:325-333: mse(UpconvBlock(init_state_low), trilinear interpolation of init_state_low to the output size
    ((48, 48, 48) for the script's 24^3 data)), fused."""
    return ic_loss(model, "trilinear")


def loss_func(output, loss_generator):
    """GS3D:334-345 on the un-padded trajectory (fused).  (The script names the function `loss_func` and the module
    instance `loss_gen`.)"""
    return loss_generator(output)
