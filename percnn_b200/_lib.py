"""ctypes binding of libpercnn_b200.so (the C-ABI declared in include/percnn_b200.h).

The product path has NO fallback: if the shared library is missing or a tensor is not on a CUDA
device the call fails loudly.  Build the library with `python -c "import __graft_entry__ as g; g.build()"`
(or `python -m percnn_b200.build`).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
#: PERCNN_B200_LIB selects another build of the same library (A/B timing of kernel variants); default: in-tree
LIB_PATH = os.environ.get("PERCNN_B200_LIB") or os.path.join(_HERE, "libpercnn_b200.so")

ABI_VERSION = 2
F32, F64 = 0, 1
CELL_PI, CELL_BURGERS, CELL_LO = 0, 1, 2
COEF_RAW, COEF_SIGMOID = 0, 1
FLAG_EVAL_BRANCH, FLAG_NO_TMA, FLAG_LO_C6 = 1, 2, 4

#: every symbol include/percnn_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    "percnn_abi_version", "percnn_last_error", "percnn_device_ok", "percnn_plan_create", "percnn_plan_destroy",
    "percnn_param_count", "percnn_state_elems", "percnn_workspace_bytes", "percnn_plan_uses_tma",
    "percnn_plan_launch_count", "percnn_params_load", "percnn_step_fwd", "percnn_step_fwd_range",
    "percnn_step_fwd_fused_halo", "percnn_step_bwd_fused_halo", "percnn_step_bwd",
    "percnn_param_grads_begin", "percnn_param_grads_finish", "percnn_rollout_fwd", "percnn_rollout_bwd",
    "percnn_rollout_fwd_host", "percnn_data_loss_fwd", "percnn_step_bwd_loss", "percnn_rollout_bwd_loss",
    "percnn_phys_loss_workspace_bytes", "percnn_phys_loss_fwd", "percnn_phys_loss_bwd",
    "percnn_slab_rollout_fwd", "percnn_slab_rollout_tape", "percnn_slab_rollout_bwd", "percnn_plan_slab_persistent", "percnn_plan_uses_tile2d", "percnn_step_rk4",
    "percnn_upscaler_sizes", "percnn_upscaler_fwd", "percnn_upscaler_bwd", "percnn_mse_workspace_bytes", "percnn_mse_fwd",
    "percnn_mse_bwd", "percnn_slab_rollout_fwd_blocked", "percnn_library_points", "percnn_library_terms", "percnn_library_theta",
)


class Desc(ctypes.Structure):
    """percnn_desc_t"""
    _fields_ = [
        ("abi_version", c_int32), ("ndim", c_int32), ("extent", c_int64 * 3), ("dtype", c_int32), ("cell", c_int32),
        ("ksize", c_int32), ("hidden", c_int32), ("coef_mode", c_int32), ("flags", c_int32), ("mu_up", c_double),
        ("dt", c_double), ("dx", c_double), ("device", c_int32), ("slab_ghost", c_int32),
    ]


class SlabLink(ctypes.Structure):
    """percnn_slab_link_t"""
    _fields_ = [
        ("peer_lo_out", c_void_p), ("peer_hi_out", c_void_p), ("my_flags", c_void_p), ("peer_lo_flags", c_void_p),
        ("peer_hi_flags", c_void_p), ("scratch", c_void_p), ("epoch", ctypes.c_uint32), ("flags", ctypes.c_uint32),
        ("peer_lo_in", c_void_p), ("peer_hi_in", c_void_p),
    ]


class SlabRing(ctypes.Structure):
    """percnn_slab_ring_t"""
    _fields_ = [
        ("buf", c_void_p * 2), ("peer_lo_buf", c_void_p * 2), ("peer_hi_buf", c_void_p * 2), ("my_flags", c_void_p),
        ("peer_lo_flags", c_void_p), ("peer_hi_flags", c_void_p), ("scratch", c_void_p),
    ]


class SlabWide(ctypes.Structure):
    """percnn_slab_wide_t"""
    _fields_ = [("buf", c_void_p * 4), ("peer_lo_buf", c_void_p * 4), ("peer_hi_buf", c_void_p * 4), ("k", c_int32),
                ("reserved", c_int32)]


class DataLoss(ctypes.Structure):
    """percnn_data_loss_t"""
    _fields_ = [
        ("target", c_void_p), ("sel", POINTER(c_uint8)), ("stride", c_int32), ("reserved", c_int32),
        ("n_total", c_int64), ("gscale", c_void_p),
    ]


class PhysLoss(ctypes.Structure):
    """percnn_phys_loss_t"""
    _fields_ = [
        ("ndim", c_int32), ("dtype", c_int32), ("extent", c_int64 * 3), ("nframes", c_int32), ("device", c_int32),
        ("diff", c_double * 2), ("poly", (c_double * 10) * 2), ("dt", c_double), ("dx", c_double),
    ]


class Upscaler(ctypes.Structure):
    """percnn_upscaler_t"""
    _fields_ = [
        ("ndim", c_int32), ("dtype", c_int32), ("channels", c_int32), ("act", c_int32), ("layers", c_int32),
        ("stride2", c_int32), ("device", c_int32), ("reserved", c_int32), ("low_extent", c_int64 * 3),
        ("out_z0", c_int64), ("out_nz", c_int64), ("out_field_stride", c_int64),
    ]


class Library(ctypes.Structure):
    """percnn_library_t"""
    _fields_ = [
        ("dtype", c_int32), ("kind", c_int32), ("H", c_int64), ("W", c_int64), ("nframes", c_int32), ("device", c_int32),
        ("dt", c_double), ("dx", c_double),
    ]


class PercnnError(RuntimeError):
    pass


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. percnn_b200 has no CPU or eager fallback; "
            "run `python -c \"import __graft_entry__ as g; g.build()\"` from the repository root.")
    L = ctypes.CDLL(LIB_PATH)
    vp = c_void_p
    L.percnn_abi_version.restype = c_int
    L.percnn_last_error.restype = c_char_p
    L.percnn_device_ok.argtypes = [c_int]
    L.percnn_plan_create.argtypes = [POINTER(Desc), POINTER(vp)]
    L.percnn_plan_destroy.argtypes = [vp]
    L.percnn_param_count.argtypes = [vp]
    L.percnn_param_count.restype = c_int64
    L.percnn_state_elems.argtypes = [vp]
    L.percnn_state_elems.restype = c_int64
    L.percnn_workspace_bytes.argtypes = [vp, c_int]
    L.percnn_workspace_bytes.restype = c_size_t
    L.percnn_plan_uses_tma.argtypes = [vp]
    L.percnn_plan_slab_persistent.argtypes = [vp]
    L.percnn_plan_uses_tile2d.argtypes = [vp]
    L.percnn_plan_launch_count.argtypes = [vp]
    L.percnn_plan_launch_count.restype = c_int64
    L.percnn_params_load.argtypes = [vp, vp, vp]
    L.percnn_step_fwd.argtypes = [vp, vp, vp, vp]
    L.percnn_step_rk4.argtypes = [vp, vp, vp, vp, vp]
    L.percnn_step_fwd_range.argtypes = [vp, vp, vp, c_int, c_int, vp]
    L.percnn_step_fwd_fused_halo.argtypes = [vp, vp, vp, POINTER(SlabLink), vp]
    L.percnn_step_bwd_fused_halo.argtypes = [vp, vp, vp, vp, vp, vp, POINTER(SlabLink), vp]
    L.percnn_step_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.percnn_param_grads_begin.argtypes = [vp, vp, vp]
    L.percnn_param_grads_finish.argtypes = [vp, vp, vp, vp, vp]
    L.percnn_rollout_fwd.argtypes = [vp, vp, vp, POINTER(c_uint8), c_int, vp, vp, vp, vp]
    L.percnn_rollout_bwd.argtypes = [vp, vp, vp, vp, POINTER(c_uint8), c_int, vp, vp, vp, vp]
    L.percnn_rollout_fwd_host.argtypes = [vp, vp, vp, vp, POINTER(c_uint8), c_int, vp]
    L.percnn_data_loss_fwd.argtypes = [vp, vp, c_int, POINTER(DataLoss), vp, vp, vp]
    L.percnn_step_bwd_loss.argtypes = [vp, vp, vp, vp, vp, c_int, c_int64, vp, vp, vp, POINTER(SlabLink), vp]
    L.percnn_rollout_bwd_loss.argtypes = [vp, vp, vp, vp, POINTER(c_uint8), POINTER(DataLoss), c_int, vp, vp, vp, vp]
    L.percnn_slab_rollout_fwd.argtypes = [vp, POINTER(SlabRing), c_int, c_int, ctypes.c_uint32, vp]
    L.percnn_slab_rollout_fwd_blocked.argtypes = [vp, POINTER(SlabRing), POINTER(SlabWide), c_int, c_int, ctypes.c_uint32, vp]
    L.percnn_slab_rollout_tape.argtypes = [vp, vp, vp, vp, POINTER(SlabRing), c_int, ctypes.c_uint32, vp]
    L.percnn_slab_rollout_bwd.argtypes = [vp, vp, vp, POINTER(DataLoss), POINTER(SlabRing), c_int, ctypes.c_uint32, vp, vp]
    L.percnn_phys_loss_workspace_bytes.restype = c_size_t
    L.percnn_phys_loss_fwd.argtypes = [POINTER(PhysLoss), vp, vp, vp, vp, vp]
    L.percnn_phys_loss_bwd.argtypes = [POINTER(PhysLoss), vp, vp, vp, vp, vp]
    L.percnn_upscaler_sizes.argtypes = [POINTER(Upscaler), POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), POINTER(c_size_t)]
    L.percnn_upscaler_fwd.argtypes = [POINTER(Upscaler), vp, vp, vp, vp, vp, vp]
    L.percnn_upscaler_bwd.argtypes = [POINTER(Upscaler), vp, vp, vp, vp, vp, c_int, vp, vp]
    L.percnn_mse_workspace_bytes.restype = c_size_t
    L.percnn_mse_fwd.argtypes = [c_int, c_int, vp, vp, c_int64, vp, vp, vp]
    L.percnn_mse_bwd.argtypes = [c_int, c_int, vp, vp, c_int64, vp, vp, c_int, vp]
    L.percnn_library_points.restype = c_int64
    L.percnn_library_points.argtypes = [POINTER(Library)]
    L.percnn_library_terms.argtypes = [POINTER(Library), vp, vp, vp]
    L.percnn_library_theta.argtypes = [POINTER(Library), vp, vp, c_int64, vp, vp, vp]
    if L.percnn_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {L.percnn_abi_version()} != {ABI_VERSION}; rebuild the library")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().percnn_last_error()
        raise PercnnError(f"libpercnn_b200 error {rc}: {msg.decode() if msg else '?'}")
