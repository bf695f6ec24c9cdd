"""ctypes binding of ``libscico_b200_xray.so`` (the C ABI in ``include/scico_b200_xray.h``).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a
compute entry point is called, an exception is raised.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libscico_b200_xray.so"
# SCICO_B200_LIB: another build of the same library (kernel A/B runs, tools/); the product loads the in-tree one
LIB_PATH = os.environ.get("SCICO_B200_LIB") or os.path.join(_HERE, LIB_NAME)

XCT_OK = 0
XCT_ERR_INVALID = -1
XCT_ERR_CUDA = -2
XCT_ERR_UNSUPPORTED = -3
XCT_ERR_NO_DEVICE = -4

SPLIT_ADMM, SPLIT_LADMM, SPLIT_PADMM = 0, 1, 2

FLAG_FORCE_GENERAL = 0x1
FLAG_NO_WALK = 0x2
FLAG_NO_HOST_PIPELINE = 0x4
FLAG_NO_JOINT = 0x8
FLAG_NO_BRICK = 0x20
FLAG_NO_TILE = 0x40
FLAG_2D_PER_CLASS = 0x80
FLAG_NO_TMA = 0x10
FLAG_NO_ADJ_VEC = 0x100
KERNEL_NAMES = {0: "general", 1: "plane", 2: "walk", 3: "brick"}

PATH_NAMES = {1: "2d_plane", 2: "2d_general", 3: "3d_sep", 4: "3d_general"}

# Every symbol include/scico_b200_xray.h declares (checked by tests/test_abi.py).
EXPORTED_SYMBOLS = (
    "xct_version",
    "xct_last_error",
    "xct_device_count",
    "xct2d_plan_create",
    "xct3d_plan_create",
    "xct_plan_destroy",
    "xct_plan_get_info",
    "xct_plan_get_classes",
    "xct2d_plan_analyse",
    "xct3d_plan_analyse",
    "xct_forward",
    "xct_adjoint",
    "xct_forward_host",
    "xct_adjoint_host",
    "xct_forward_host_async",
    "xct_adjoint_host_async",
    "xct_host_wait",
    "xct3d_debug_weights",
    "xct2d_debug_weights",
    "xct_adjoint_scatter",
    "xct_peer_alloc",
    "xct_peer_open",
    "xct_peer_zero",
    "xct_peer_copy_out",
    "xct_sum_slots",
    "xct_peer_signal",
    "xct_peer_wait",
    "xct_op_register_2d",
    "xct_op_register_3d",
    "xct_op_retain",
    "xct_op_release",
    "xct_op_plan",
    "xct_op_apply",
    "xct_peer_close",
    "xct_peer_free",
    "xct_launch_count",
    "xct_launch_count_reset",
    "xct_tv_primal_step",
    "xct_tv_dual_step",
    "xct_l2_dual_step",
    "xct_tv_primal_step_stat",
    "xct_tv_dual_step_stat",
    "xct_l2_dual_step_stat",
    "xct_tv_norm",
    "xct_grad_prox_step_stat",
    "xct_sino_prox_step_stat",
    "xct_fd_forward",
    "xct_fd_adjoint",
    "xct_grad_prox_step",
    "xct_sino_prox_step",
    "xct_grad_primal_step",
    "xct_admm_rhs",
    "xct_cg_init",
    "xct_cg_lhs",
    "xct_cg_update_xr",
    "xct_cg_update_p",
)


class Geom2D(ctypes.Structure):
    _fields_ = [
        ("n0", c_int32),
        ("n1", c_int32),
        ("num_views", c_int32),
        ("det_count", c_int32),
        ("view_table", POINTER(c_float)),
        ("device", c_int32),
        ("flags", c_uint32),
    ]


class Geom3D(ctypes.Structure):
    _fields_ = [
        ("n0", c_int32),
        ("n1", c_int32),
        ("n2", c_int32),
        ("d0", c_int32),
        ("d1", c_int32),
        ("num_views", c_int32),
        ("matrices", POINTER(c_float)),
        ("slice_offset", c_int32),
        ("det_row_offset", c_int32),
        ("det_rows_total", c_int32),
        ("device", c_int32),
        ("flags", c_uint32),
    ]


class PlanInfo(ctypes.Structure):
    _fields_ = [
        ("ndim", c_int32),
        ("path", c_int32),
        ("num_views", c_int32),
        ("fwd_lane_stride", c_int32),
        ("row_aligned", c_int32),
        ("device", c_int32),
        ("adj_kernel", c_int32),
        ("fwd_kernel", c_int32),
        ("fwd_joint", c_int32),
        ("adj_tma", c_int32),
        ("in_elems", c_int64),
        ("out_elems", c_int64),
        ("updates", c_int64),
    ]


class PlanClasses(ctypes.Structure):
    _fields_ = [
        ("joint_views", c_int32 * 8),
        ("two_bin_views", c_int32 * 4),
        ("adj_jump_views", c_int32),
        ("rows_unit", c_int32),
        ("rows_consecutive", c_int32),
        ("fwd_cold", c_int32),
        ("brick_views", c_int32 * 6),
        ("fwd_tile", c_int32),
        ("adj_interleaved", c_int32),
    ]


MAX_ROUTE_PARTS = 16


class OutRoute(ctypes.Structure):
    _fields_ = [
        ("nparts", c_int32),
        ("row_begin", c_int32 * (MAX_ROUTE_PARTS + 1)),
        ("ptr", c_void_p * MAX_ROUTE_PARTS),
        ("store", c_int32),
    ]


class IpcHandle(ctypes.Structure):
    _fields_ = [("bytes", ctypes.c_ubyte * 64)]


class TvBlock(ctypes.Structure):
    _fields_ = [("n0", c_int32), ("n1", c_int32), ("n2", c_int32), ("is_first", c_int32), ("is_last", c_int32)]


class XctError(RuntimeError):
    """Raised when a C-ABI call returns a negative status."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"scico_b200 native call failed ({code}): {msg}")
        self.code = code


_lib = None


def lib() -> ctypes.CDLL:
    """Load the native library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or scico_b200/build.py).  scico_b200 has no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    L.xct_version.restype = ctypes.c_int
    L.xct_last_error.restype = c_char_p
    L.xct_device_count.restype = ctypes.c_int
    L.xct2d_plan_create.argtypes = [POINTER(c_void_p), POINTER(Geom2D)]
    L.xct3d_plan_create.argtypes = [POINTER(c_void_p), POINTER(Geom3D)]
    L.xct_plan_destroy.argtypes = [c_void_p]
    L.xct_plan_destroy.restype = None
    L.xct_plan_get_info.argtypes = [c_void_p, POINTER(PlanInfo)]
    L.xct_plan_get_classes.argtypes = [c_void_p, POINTER(PlanClasses)]
    L.xct2d_plan_analyse.argtypes = [POINTER(Geom2D), POINTER(PlanInfo), POINTER(PlanClasses)]
    L.xct3d_plan_analyse.argtypes = [POINTER(Geom3D), POINTER(PlanInfo), POINTER(PlanClasses)]
    for name in ("xct_forward", "xct_adjoint"):
        getattr(L, name).argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    for name in ("xct_forward_host", "xct_adjoint_host", "xct_forward_host_async", "xct_adjoint_host_async"):
        getattr(L, name).argtypes = [c_void_p, c_void_p, c_void_p, c_int32]
    L.xct_host_wait.argtypes = [c_void_p]
    L.xct_adjoint_scatter.argtypes = [c_void_p, c_void_p, POINTER(OutRoute), c_void_p]
    L.xct_peer_alloc.argtypes = [c_int32, ctypes.c_size_t, POINTER(c_void_p), POINTER(IpcHandle)]
    L.xct_peer_open.argtypes = [c_int32, POINTER(IpcHandle), POINTER(c_void_p)]
    L.xct_peer_zero.argtypes = [c_int32, c_void_p, ctypes.c_size_t, c_void_p]
    L.xct_peer_copy_out.argtypes = [c_int32, c_void_p, c_void_p, ctypes.c_size_t, c_void_p]
    L.xct_sum_slots.argtypes = [c_int32, c_void_p, c_void_p, c_int32, ctypes.c_size_t, ctypes.c_size_t, c_void_p]
    L.xct_peer_signal.argtypes = [c_int32, POINTER(c_void_p), c_int32, c_int32, c_void_p]
    L.xct_peer_wait.argtypes = [c_int32, c_void_p, c_int32, c_int32, ctypes.c_double, c_void_p, c_void_p]
    L.xct_op_register_2d.argtypes = [POINTER(Geom2D), POINTER(c_int64)]
    L.xct_op_register_3d.argtypes = [POINTER(Geom3D), POINTER(c_int64)]
    L.xct_op_retain.argtypes = [c_int64]
    L.xct_op_release.argtypes = [c_int64]
    L.xct_op_plan.argtypes = [c_int64, c_int32, POINTER(c_void_p)]
    L.xct_op_apply.argtypes = [c_int64, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p]
    L.xct_peer_close.argtypes = [c_int32, c_void_p]
    L.xct_peer_free.argtypes = [c_int32, c_void_p]
    L.xct3d_debug_weights.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p]
    L.xct2d_debug_weights.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p]
    cf = ctypes.c_float
    L.xct_tv_primal_step.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, cf, cf, c_int32, c_void_p]
    L.xct_tv_dual_step.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, cf, cf, c_void_p]
    L.xct_l2_dual_step.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, cf, c_void_p]
    L.xct_tv_primal_step_stat.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, cf, cf, c_int32,
                                          c_void_p, c_void_p]
    L.xct_tv_dual_step_stat.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, cf, cf, c_void_p, c_void_p]
    L.xct_l2_dual_step_stat.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, cf, c_void_p, cf, c_int64, c_int32, c_int32,
                                        c_int32, c_void_p, c_void_p]
    L.xct_tv_norm.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, c_void_p]
    L.xct_grad_prox_step_stat.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, cf, cf, cf, c_int32,
                                          c_void_p, c_void_p]
    L.xct_sino_prox_step_stat.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, cf, cf, c_int32, c_int64,
                                          c_int32, c_int32, c_int32, c_void_p, c_void_p]
    L.xct_fd_forward.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, c_void_p]
    L.xct_fd_adjoint.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, c_void_p]
    L.xct_grad_prox_step.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, cf, cf, cf, c_int32, c_void_p]
    L.xct_sino_prox_step.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, cf, cf, c_int32, c_void_p]
    L.xct_grad_primal_step.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, c_void_p, cf, cf, c_int32, c_void_p]
    L.xct_admm_rhs.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, c_void_p, cf, c_void_p, c_void_p, c_void_p]
    L.xct_cg_init.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, cf, c_void_p, c_void_p, c_void_p, c_void_p]
    L.xct_cg_lhs.argtypes = [POINTER(TvBlock), c_void_p, c_void_p, c_void_p, c_void_p, cf, c_void_p, c_void_p, c_void_p, c_void_p]
    L.xct_cg_update_xr.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.xct_cg_update_p.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.xct_launch_count.restype = c_int64
    L.xct_launch_count_reset.restype = None
    _lib = L
    return L


def check(code: int) -> None:
    if code != XCT_OK:
        raise XctError(code, lib().xct_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().xct_launch_count())


def launch_count_reset() -> None:
    lib().xct_launch_count_reset()
