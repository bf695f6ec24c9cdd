"""Host-side geometry of the projector pair (NumPy, float32 where the reference is float32).

Per-view scalars are computed once here and handed to the kernels as a table: ``cos``/``sin``
are never evaluated on the device, so kernel and checker see bit-identical coefficients.

Reference: ``scico/linop/xray/_xray2d.py:326-347`` (2D per-view part of ``_calc_weights``) and
``scico/linop/xray/_xray3d.py:268-327`` (``matrices_from_euler_angles``).
"""

from __future__ import annotations

import numpy as np

f32 = np.float32


def view_table_2d(angles, x0, dx, y0) -> np.ndarray:
    """(V, 4) float32 table ``(Pxmin, Pdx0, Pdx1, width)``.

    Everything is float32 because the reference traces ``x0, dx, y0, angles`` into a jitted
    function with x64 disabled (``_xray2d.py:223-233``).
    """
    ang = np.asarray(angles, dtype=np.float64).astype(f32).reshape(-1)
    x0 = np.asarray(x0, dtype=np.float64).astype(f32).reshape(2)
    dx = np.asarray(dx, dtype=np.float64).astype(f32).reshape(2)
    y0 = f32(y0)
    c, s = np.cos(ang).astype(f32), np.sin(ang).astype(f32)
    px0 = (x0[0] * c + x0[1] * s) - y0
    pdx0, pdx1 = dx[0] * c, dx[1] * s
    pxmin = np.minimum(np.minimum(px0, px0 + pdx0), np.minimum(px0 + pdx1, (px0 + pdx0) + pdx1))
    d1, d2 = np.abs(pdx0 + pdx1), np.abs(pdx0 - pdx1)
    width = (np.maximum(d1, d2) + np.minimum(d1, d2)) / f32(2)
    return np.ascontiguousarray(np.stack([pxmin, pdx0, pdx1, width], axis=1), dtype=f32)


def max_projected_width(angles, dx) -> float:
    """Largest projected pixel width over all views (``_xray2d.py:96-100``)."""
    ang = np.asarray(angles, dtype=np.float64).reshape(-1)
    p0, p1 = dx[0] * np.cos(ang), dx[1] * np.sin(ang)
    return float(np.max(np.maximum(np.abs(p0 + p1), np.abs(p0 - p1))))


def _axis_rotation(axis: str, a: np.ndarray) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    R = np.zeros(a.shape + (3, 3))
    i = "xyz".index(axis)
    j, k = (i + 1) % 3, (i + 2) % 3
    R[..., i, i] = 1.0
    R[..., j, j] = c
    R[..., j, k] = -s
    R[..., k, j] = s
    R[..., k, k] = c
    return R


def euler_matrices(seq: str, angles, degrees: bool = False) -> np.ndarray:
    """Rotation matrices with ``scipy.spatial.transform.Rotation.from_euler`` semantics
    (upper-case ``seq`` = intrinsic, lower-case = extrinsic), (V, 3, 3) float64."""
    angles = np.asarray(angles, dtype=np.float64)
    if angles.ndim == 1:
        angles = angles[:, None] if len(seq) == 1 else angles[None, :]
    if angles.ndim != 2 or angles.shape[1] != len(seq) or not 1 <= len(seq) <= 3:
        raise ValueError(f"Euler sequence {seq!r} does not match angles of shape {angles.shape}.")
    intrinsic = seq.isupper()
    if not (intrinsic or seq.islower()) or any(ch not in "xyz" for ch in seq.lower()):
        raise ValueError(f"Invalid Euler sequence {seq!r}.")
    if degrees:
        angles = np.deg2rad(angles)
    R = np.broadcast_to(np.eye(3), (angles.shape[0], 3, 3)).copy()
    for n, ax in enumerate(seq.lower()):
        Rn = _axis_rotation(ax, angles[:, n])
        R = R @ Rn if intrinsic else Rn @ R
    return R


def matrices_from_euler_angles(
    input_shape, output_shape, seq, angles, degrees=False, voxel_spacing=None, det_spacing=None
) -> np.ndarray:
    """(V, 2, 4) float64 homogeneous projection matrices (``_xray3d.py:268-327``):
    rotate, drop the last row, scale by voxel / detector spacing, centre on the detector."""
    voxel_spacing = np.ones(3) if voxel_spacing is None else np.asarray(voxel_spacing, dtype=np.float64)
    det_spacing = np.ones(2) if det_spacing is None else np.asarray(det_spacing, dtype=np.float64)
    M = euler_matrices(seq, angles, degrees)[:, :2, :]
    M = M * voxel_spacing[None, None, :] / det_spacing[None, :, None]
    centre = np.asarray(input_shape, dtype=np.float64) / 2
    t = -(M @ centre) + np.asarray(output_shape, dtype=np.float64) / 2
    return np.concatenate([M, t[..., None]], axis=2)


def is_axis0_separable(matrices) -> bool:
    """True when detector rows depend on voxel axis 0 only and columns on axes 1, 2 only
    (e.g. any rotation about axis 0): the geometry z-slabs shard without communication."""
    M = np.asarray(matrices, dtype=f32)
    return bool(np.all(M[:, 0, 1] == 0) and np.all(M[:, 0, 2] == 0) and np.all(M[:, 1, 0] == 0))


def slab_row_range(matrices, z0: int, z1: int, det_rows: int) -> tuple[int, int]:
    """Detector rows ``[r0, r1)`` touched by voxel slices ``[z0, z1)`` of a separable geometry,
    using the same fp32 row arithmetic as the kernels (``_xray3d.py:216,223-225``)."""
    M = np.asarray(matrices, dtype=f32)
    xi = (np.arange(z0, z1, dtype=np.int32).astype(f32) + f32(0.5))[None, :]
    left = (M[:, 0, 0:1] * xi + M[:, 0, 3:4]) - f32(0.25)
    tn = np.minimum(np.ceil(left) - left, f32(0.5))
    lo = np.floor(left).astype(np.int64)
    first = np.where(tn > 0, lo, lo + 1)  # row lo carries weight only if to_next > 0
    last = np.where(tn < f32(0.5), lo + 1, lo)
    r0 = int(max(0, first.min()))
    r1 = int(min(det_rows, last.max() + 1))
    return r0, max(r0, r1)
