"""Drop-in for Stage-1 Burgers (BUR1:38-303): 5x5 Pi convs on the manually padded state."""
from ._stage1 import Stage1Cell, Stage1RCNN, get_ic_loss, upscaler  # noqa: F401


class RCNNCell(Stage1Cell):
    _dx, _dt, _nu_up = 1 / 100, 0.00025, 0.01       # BUR1:94-96


class RCNN(Stage1RCNN):
    cell_cls = RCNNCell
