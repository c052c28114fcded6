"""Drop-in for the classes of DataDrivenModeling/2d_gs_rd/train_2drd.py (GS2D:26-190)."""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib, losses
from ..cells import FusedRCNN, PiCell
from ..upscaler import FusedUpscaler, ic_loss


class upscaler(FusedUpscaler):
    """Low-res -> full-res initial-state generator (GS2D:26-41): two stride-2 transposed convs and a 1x1 conv.

    The layers are the reference's (same construction order, so the same initial values under a given seed, and the
    same state_dict keys `convnet.{0,2,3}.*`); they only hold the parameters.  `forward` runs the fused kernels
    (percnn_upscaler_fwd / _bwd).
    """
    up_ndim, up_channels, up_act, up_stride2 = 2, 8, "sigmoid", 2

    def __init__(self):
        super().__init__()
        self.layers = [
            nn.ConvTranspose2d(2, 8, kernel_size=5, padding=2, stride=2, output_padding=1, bias=True),
            nn.Sigmoid(),
            nn.ConvTranspose2d(8, 8, kernel_size=5, padding=2, stride=2, output_padding=1, bias=True),
            nn.Conv2d(8, 2, 1, 1, padding=0, bias=True),
        ]
        self.convnet = nn.Sequential(*self.layers)

    def _up_modules(self):
        return [self.convnet[0], self.convnet[2], self.convnet[3]]


class RCNNCell(PiCell):
    """GS2D:43-121: fp32, 1x1 Pi convs, alpha = mu_up * sigmoid(CA)."""

    def __init__(self, input_channels, hidden_channels, input_kernel_size):
        super().__init__()
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.input_kernel_size = 5       # GS2D:53 ignores the argument
        self.input_stride = 1
        self.mu_up = 3.99e-5
        np.random.seed(1234)             # GS2D:60 -- constructor side effect kept on purpose
        ca, cb = (np.random.rand() - 0.5) * 2, (np.random.rand() - 0.5) * 2
        self._build(ndim=2, dtype=torch.float32, ksize=1, hidden=hidden_channels, dx=0.01, dt=0.5,
                    coef_mode=_lib.COEF_SIGMOID, mu_up=self.mu_up, coef_names=("CA", "CB"), coef_init=(ca, cb),
                    init_scale=0.02, init_kind="xavier")


class RCNN(FusedRCNN):
    def __init__(self, input_channels, hidden_channels, init_state_low, input_kernel_size, step=1, effective_step=[1]):
        super().__init__()
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.input_kernel_size = input_kernel_size
        self.init_state_low = init_state_low
        self.init_state = []
        self.UpconvBlock = upscaler()
        self._setup(RCNNCell(input_channels=input_channels, hidden_channels=hidden_channels,
                             input_kernel_size=input_kernel_size), step, effective_step)


class loss_generator(losses.LossGenerator):
    """GS2D:241-330 `loss_generator(dt, dx)`: Gray-Scott residual with Du = 2e-5, Dv = Du/4, f = 1/25, k = 3/50."""

    def __init__(self, dt=(1.0 / 2), dx=(1.0 / 100)):
        super().__init__(losses.gray_scott_spec(2e-5, 2e-5 / 4, 1 / 25, 3 / 50, dt, dx))


def get_ic_loss(model):
    """GS2D:331-338: mse(UpconvBlock(init_state_low), bicubic interpolation of init_state_low to the output size
    ((100, 100) for the script's 25 x 25 data)), fused (percnn_mse_fwd / _bwd + the upscaler kernels)."""
    return ic_loss(model, "bicubic")


def loss_gen(output, loss_func):
    """GS2D:340-353 on the un-padded trajectory (fused)."""
    return loss_func(output)
