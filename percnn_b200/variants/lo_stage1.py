"""Drop-in for Stage-1 lambda-omega (LO1:38-296): same cell with conv-internal circular padding, which is
the same arithmetic as BUR1's manual padding."""
from ._stage1 import Stage1Cell, Stage1RCNN, get_ic_loss, upscaler  # noqa: F401


class RCNNCell(Stage1Cell):
    _dx, _dt, _nu_up = 0.2, 0.0125, 0.2              # LO1:93-95

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.input_padding = self.input_kernel_size // 2


class RCNN(Stage1RCNN):
    cell_cls = RCNNCell
