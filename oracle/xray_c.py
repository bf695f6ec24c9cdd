"""ctypes wrapper over ``libxray_oracle.so`` (C restatement; TEST INFRASTRUCTURE)."""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libxray_oracle.so")
    src = os.path.join(_HERE, "xray_c.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libxray_oracle.so"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.xo_num_threads.restype = ctypes.c_int
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_fp)


def num_threads() -> int:
    return int(lib().xo_num_threads())


def set_num_threads(n: int) -> None:
    lib().xo_set_num_threads(int(n))


def project_2d(im, table, ny, fused=False):
    im, pim = _f(im)
    table, pt = _f(table)
    V = table.shape[0]
    out = np.empty((V, ny), dtype=np.float32)
    lib().xo_project_2d(pim, pt, V, im.shape[0], im.shape[1], int(ny), out.ctypes.data_as(_fp), int(fused))
    return out


def back_project_2d(y, table, nx):
    y, py = _f(y)
    table, pt = _f(table)
    out = np.empty(tuple(nx), dtype=np.float32)
    lib().xo_back_project_2d(py, pt, y.shape[0], int(nx[0]), int(nx[1]), y.shape[1], out.ctypes.data_as(_fp))
    return out


def weights_2d(table_row, nx):
    t, pt = _f(table_row)
    inds = np.empty(tuple(nx), dtype=np.int32)
    w = np.empty(tuple(nx), dtype=np.float32)
    lib().xo_weights_2d(pt, int(nx[0]), int(nx[1]), inds.ctypes.data_as(_ip), w.ctypes.data_as(_fp))
    return inds, w


def project_3d(im, matrices, det_shape, slice_offset=0, fused=False):
    im, pim = _f(im)
    M, pm = _f(matrices)
    V = M.shape[0]
    out = np.empty((V, int(det_shape[0]), int(det_shape[1])), dtype=np.float32)
    lib().xo_project_3d(pim, pm, V, *map(int, im.shape), int(det_shape[0]), int(det_shape[1]),
                        int(slice_offset), out.ctypes.data_as(_fp), int(fused))
    return out


def back_project_3d(proj, matrices, input_shape, slice_offset=0):
    proj, pp = _f(proj)
    M, pm = _f(matrices)
    out = np.empty(tuple(input_shape), dtype=np.float32)
    lib().xo_back_project_3d(pp, pm, M.shape[0], *map(int, input_shape), proj.shape[1], proj.shape[2],
                             int(slice_offset), out.ctypes.data_as(_fp))
    return out


def weights_3d(matrix, input_shape, det_shape, slice_offset=0):
    M, pm = _f(matrix)
    n = tuple(map(int, input_shape))
    ul = np.empty((2,) + n, dtype=np.int32)
    w = np.empty((4,) + n, dtype=np.float32)
    lib().xo_weights_3d(pm, *n, int(det_shape[0]), int(det_shape[1]), int(slice_offset),
                        ul.ctypes.data_as(_ip), w.ctypes.data_as(_fp))
    return ul, w


def back_project_3d_points(proj, matrices, points, slice_offset=0):
    """Back projection at selected voxels ``points`` (npts, 3) int32 -> (npts,) float32."""
    proj, pp = _f(proj)
    M, pm = _f(matrices)
    pts = np.ascontiguousarray(points, dtype=np.int32)
    out = np.empty(len(pts), dtype=np.float32)
    lib().xo_back_project_3d_points(pp, pm, M.shape[0], proj.shape[1], proj.shape[2], int(slice_offset),
                                    pts.ctypes.data_as(_ip), len(pts), out.ctypes.data_as(_fp))
    return out
