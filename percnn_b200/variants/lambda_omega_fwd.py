"""Drop-in for the classes of ForwardSimulationOfPDEs/2d_lambda_omega/percnn_LO_eqn.py (FWD:24-218)."""
import torch
import torch.nn as nn

from .. import _lib, losses
from ..cells import FusedRCNN, PiCell


class RCNNCell(PiCell):
    """FWD:24-112: fp64, 1x1 Pi convs with 4 channels, raw trainable diffusion coefficients DA, DB."""

    def __init__(self, input_kernel_size=1, input_stride=1, input_padding=0):
        super().__init__()
        if input_kernel_size != 1 or input_stride != 1:
            raise ValueError("the fused lambda-omega cell supports input_kernel_size=1, input_stride=1 (as FWD:499 uses)")
        self.input_kernel_size = input_kernel_size
        self.input_stride = input_stride
        self.input_padding = input_padding
        self._build(ndim=2, dtype=torch.float64, ksize=1, hidden=4, dx=0.2, dt=0.0125, coef_mode=_lib.COEF_RAW,
                    mu_up=1.0, coef_names=("DA", "DB"), coef_init=(0.2, 0.2), init_scale=0.5, init_kind="uniform")


class RCNN(FusedRCNN):
    """FWD:124-218: the initial state is given (no upscaler) and the cell is registered as `rcnn_cell`."""

    cell_attr = "rcnn_cell"

    def __init__(self, input_kernel_size, ini_state, input_stride, input_padding, step=1, effective_step=[1]):
        super().__init__()
        self.input_kernel_size = input_kernel_size
        self.input_stride = input_stride
        self.input_padding = input_padding
        self.init_state = torch.as_tensor(ini_state, dtype=torch.float64)
        self._setup(RCNNCell(input_kernel_size=input_kernel_size, input_stride=input_stride,
                             input_padding=input_padding), step, effective_step)

    def _initial_state(self):
        return self.init_state

    def load_state_dict(self, state_dict, strict=True, **kw):
        # the shipped rcnn_pde.pt uses the prefix `crnn_cell.` while FWD:160 registers `rcnn_cell`
        fixed = {(k.replace("crnn_cell.", "rcnn_cell.", 1) if k.startswith("crnn_cell.") else k): v
                 for k, v in state_dict.items()}
        return super().load_state_dict(fixed, strict=strict, **kw)


class loss_generator(losses.LossGenerator):
    """FWD:265-342 `loss_generator(dt, dx)`: lambda-omega residual, 0.1 Lap + analytic reaction - d/dt."""

    def __init__(self, dt=0.0125, dx=0.2):
        super().__init__(losses.lambda_omega_spec(dt, dx))


def loss_gen(output, loss_func):
    """FWD:344-357: takes the UN-padded trajectory `torch.cat(outputs)`; padding, Laplacian, time difference,
    reaction term and both MSEs happen inside one fused kernel (percnn_phys_loss_fwd)."""
    return loss_func(output)
