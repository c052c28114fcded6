"""Drop-in for Stage-3 lambda-omega (LO3:54-353): f = nu Lap + cubic polynomial in (u, v)."""
import torch

from .. import _lib
from ..cells import Conv2dDerivative, PhysicsCell, laplace_table
from ..engine import CellSpec
from ._stage3 import Stage3RCNN, _scalar, get_ic_loss, upscaler  # noqa: F401

_INIT = (("nu_u", 0.09465), ("nu_v", 0.09455), ("C1_u", 1.0081), ("C2_u", -1.0167), ("C3_u", 0.9973),
         ("C4_u", -1.0176), ("C5_u", 0.9981), ("C1_v", 0.9873), ("C2_v", -0.9987), ("C3_v", -0.9945),
         ("C4_v", -0.9985), ("C5_v", -0.9928))


class RCNNCell(PhysicsCell):
    """LO3:83-215.  `with_c6=True` adds the coefficient C6_v of the 10 %-noise twin script."""

    with_c6 = False

    def __init__(self, input_channels, hidden_channels, output_channels, input_kernel_size, input_stride, input_padding):
        super().__init__()
        self.ndim, self.dtype = 2, torch.float64
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.output_channels = output_channels
        self.input_kernel_size = 5
        self.input_stride = input_stride
        self.input_padding = 2
        for name, val in _INIT:
            setattr(self, name, _scalar(val))
        if self.with_c6:
            self.C6_v = _scalar(0.0065)
        self.dx = self.dy = 0.2
        self.dt = 0.0125
        self.laplace_op = Conv2dDerivative(laplace_table(2).tolist(), self.dx ** 2, 5, "laplace_operator")

    def _spec(self):
        return CellSpec(cell=_lib.CELL_LO, ndim=2, dtype=self.dtype, ksize=0, hidden=0, coef_mode=_lib.COEF_RAW,
                        mu_up=1.0, dt=float(self.dt), dx=float(self.dx), flags=_lib.FLAG_LO_C6 if self.with_c6 else 0)

    def f_rhs(self, u, v):
        f_u = (self.nu_u * self.laplace_op(u) + self.C1_u * u + self.C2_u * u ** 3 + self.C3_u * u ** 2 * v
               + self.C4_u * u * v ** 2 + self.C5_u * v ** 3)
        f_v = (self.nu_v * self.laplace_op(v) + self.C1_v * v + self.C2_v * u ** 3 + self.C3_v * u ** 2 * v
               + self.C4_v * u * v ** 2 + self.C5_v * v ** 3)
        if self.with_c6:
            f_v = f_v + self.C6_v * u
        return f_u, f_v

    def show_coef(self):
        from prettytable import PrettyTable
        names = [n for n, _ in _INIT] + (["C6_v"] if self.with_c6 else [])
        table = PrettyTable()
        table.field_names = ["\\"] + names
        table.add_row(["Identified"] + [getattr(self, n).item() for n in names])
        print(table)


class RCNNCellNoisy(RCNNCell):
    with_c6 = True


class RCNN(Stage3RCNN):
    cell_cls = RCNNCell


class RCNNNoisy(Stage3RCNN):
    cell_cls = RCNNCellNoisy
