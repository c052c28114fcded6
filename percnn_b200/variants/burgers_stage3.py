"""Drop-in for Stage-3 Burgers (BUR3:54-356): f = nu Lap + C1 u d/dx + C2 v d/dy, the advection-stencil cell."""
import torch

from .. import _lib
from ..cells import Conv2dDerivative, PhysicsCell, derivative_table, laplace_table
from ..engine import CellSpec
from ._stage3 import Stage3RCNN, _scalar, get_ic_loss, upscaler  # noqa: F401


class RCNNCell(PhysicsCell):
    def __init__(self, input_channels, hidden_channels, output_channels, input_kernel_size, input_stride, input_padding):
        super().__init__()
        self.ndim, self.dtype = 2, torch.float64
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.output_channels = output_channels
        self.input_kernel_size = 5
        self.input_stride = input_stride
        self.input_padding = 2
        # identified Stage-2 coefficients the script starts from (BUR3:123-130)
        for name, val in (("nu_u", 0.0050078), ("nu_v", 0.0050228), ("C1_u", -0.982252), ("C2_u", -0.992132),
                          ("C1_v", -0.983758), ("C2_v", -0.971269)):
            setattr(self, name, _scalar(val))
        self.dx = self.dy = 1 / 100
        self.dt = 0.00025
        self.laplace_op = Conv2dDerivative(laplace_table(2).tolist(), self.dx ** 2, 5, "laplace_operator")
        self.dx_op = Conv2dDerivative(derivative_table(0).tolist(), self.dx, 5, "dx_operator")
        self.dy_op = Conv2dDerivative(derivative_table(1).tolist(), self.dy, 5, "dy_operator")

    def _spec(self):
        return CellSpec(cell=_lib.CELL_BURGERS, ndim=2, dtype=self.dtype, ksize=0, hidden=0, coef_mode=_lib.COEF_RAW,
                        mu_up=1.0, dt=float(self.dt), dx=float(self.dx))

    def f_rhs(self, u, v):
        f_u = self.nu_u * self.laplace_op(u) + self.C1_u * u * self.dx_op(u) + self.C2_u * v * self.dy_op(u)
        f_v = self.nu_v * self.laplace_op(v) + self.C1_v * u * self.dx_op(v) + self.C2_v * v * self.dy_op(v)
        return f_u, f_v

    def show_coef(self):
        from prettytable import PrettyTable
        table = PrettyTable()
        table.field_names = ["\\", "nu_u", "nu_v", "Cu_1", "Cu_2", "Cv_1", "Cv_2"]
        table.add_row(["True", 0.005, 0.005, -1, -1, -1, -1])
        table.add_row(["Identified"] + [getattr(self, n).item() for n in ("nu_u", "nu_v", "C1_u", "C2_u", "C1_v", "C2_v")])
        print(table)


class RCNN(Stage3RCNN):
    cell_cls = RCNNCell
