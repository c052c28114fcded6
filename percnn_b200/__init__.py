"""percnn_b200 -- B200-native fused recurrent cell for PeRCNN (isds-neu/PeRCNN).

Public surface:
    percnn_b200.variants.<script family>.{RCNNCell, RCNN, upscaler, loss_generator, loss_gen, get_ic_loss}
                                                                       drop-in nn.Modules / functions
    percnn_b200.engine.{CellSpec, Plan, get_plan, rollout_states, rollout_emit}
    percnn_b200.losses, percnn_b200.upscaler, percnn_b200.library      fused physics loss, initial-state generator + IC loss,
                                                                       Stage-2 library of candidate terms
    percnn_b200.halo                                                   slab decomposition over N GPUs
    include/percnn_b200.h + libpercnn_b200.so                          the C-ABI underneath

Importing the package does not need a GPU; using a cell does (no CPU fallback).
"""
from . import _lib  # noqa: F401
from .engine import CellSpec, Plan, get_plan, rollout_emit, rollout_states  # noqa: F401

__all__ = ["CellSpec", "Plan", "get_plan", "rollout_emit", "rollout_states"]
