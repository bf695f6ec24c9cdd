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
    try:  # the reference's own call (_xray3d.py:304): bit-identical matrices when scipy is there
        from scipy.spatial.transform import Rotation

        return np.asarray(Rotation.from_euler(seq, angles, degrees=degrees).as_matrix(),
                          dtype=np.float64).reshape(-1, 3, 3)
    except ImportError:
        pass
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
    # the reference's own NumPy expressions (_xray3d.py:308-325), so that every entry rounds the same way
    M = np.einsum("vmn,nn->vmn", M, np.diag(voxel_spacing))
    M = np.einsum("mm,vmn->vmn", np.diag(1 / det_spacing), M)
    centre = np.array(input_shape) / 2
    t = -np.einsum("vmn,n->vm", M, centre) + np.array(output_shape) / 2
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


# ---------------------------------------------------------------------------------------------
# ASTRA "parallel3d_vec" <-> projection matrices, without ASTRA
# ---------------------------------------------------------------------------------------------
# Conventions (scico/linop/xray/astra/_astra_3d.py:185-307,595-631 and the ASTRA geometry docs):
# a 12-vector per view holds (ray, d, u, v) in WORLD (x, y, z) order: ray direction, detector centre,
# step from detector pixel (0,0) to (0,1) [columns], step from (0,0) to (1,0) [rows].  The volume
# array is indexed (z, y, x) = (axis 0, 1, 2), has unit voxels and is centred on the world origin
# (astra.create_vol_geom: window [-n/2, n/2] per axis); index i sits at world coordinate
# i - (n/2 - 1/2).  Detector pixel indices are (row, col), pixel (r, c) centred at
# d + (c - (C-1)/2) u + (r - (R-1)/2) v.


def angle_to_vector(det_spacing, angles) -> np.ndarray:
    """ASTRA "parallel3d" (det_spacing, angles) as "parallel3d_vec" vectors, shape (V, 12)
    (``_astra_3d.py:595-612``): rotation about the world z axis, rays in the x-y plane."""
    a = np.asarray(angles, dtype=np.float64).reshape(-1)
    s, c = np.sin(a), np.cos(a)
    vec = np.zeros((a.size, 12))
    vec[:, 0], vec[:, 1] = s, -c                                    # ray
    vec[:, 6], vec[:, 7] = c * det_spacing[0], s * det_spacing[0]   # u: along detector columns
    vec[:, 11] = det_spacing[1]                                     # v: along detector rows (world z)
    return vec


def rotate_vectors(vectors, rot) -> np.ndarray:
    """Rotate every 3-vector of "parallel3d_vec" vectors (``_astra_3d.py:615-631``).  `rot`: anything
    with an ``apply`` method (``scipy.spatial.transform.Rotation``) or a (3, 3) matrix."""
    v = np.array(vectors, dtype=np.float64, copy=True).reshape(-1, 4, 3)
    if hasattr(rot, "apply"):
        out = np.stack([rot.apply(v[:, k]) for k in range(4)], axis=1)
    else:
        out = v @ np.asarray(rot, dtype=np.float64).T
    return out.reshape(-1, 12)


def volume_coords_to_world_coords(idx, input_shape) -> np.ndarray:
    """Index coordinates (..., 3) in (axis 0, 1, 2) order -> world (x, y, z) for a unit-voxel volume
    of shape `input_shape` centred on the origin (``_astra_3d.py:118-182``)."""
    n = np.asarray(input_shape, dtype=np.float64)[::-1]
    return np.asarray(idx, dtype=np.float64)[..., ::-1] - (n / 2 - 0.5)


def project_world_coordinates(x, ray, d, u, v, det_shape) -> np.ndarray:
    """World points (..., 3) -> detector index coordinates (..., 2) as (row, col): express x - d in
    the basis (ray, u, v), drop the ray component (``_astra_3d.py:86-115``)."""
    basis = np.stack([np.asarray(ray, float), np.asarray(u, float), np.asarray(v, float)], axis=1)
    coef = (np.asarray(x, dtype=np.float64) - np.asarray(d, float)) @ np.linalg.pinv(basis).T
    rows = coef[..., 2] + (det_shape[0] / 2 - 0.5)
    cols = coef[..., 1] + (det_shape[1] / 2 - 0.5)
    return np.stack([rows, cols], axis=-1)


def convert_to_scico_geometry(input_shape, det_count, det_spacing=None, angles=None, vectors=None) -> np.ndarray:
    """(V, 2, 4) projection matrices for an ASTRA-style geometry given either (`det_spacing`,
    `angles`) ["parallel3d"] or `vectors` ["parallel3d_vec"] (``_astra_3d.py:265-307``).

    The map index -> detector index is affine, so its matrix is read off the images of the index
    origin and the three unit steps (the reference solves the same 4-point system per view)."""
    if angles is not None and vectors is not None:
        raise ValueError("Arguments 'angles' and 'vectors' are mutually exclusive.")
    if angles is None and vectors is None:
        raise ValueError("Exactly one of arguments 'angles' and 'vectors' must be provided.")
    if vectors is None:
        if det_spacing is None:
            raise ValueError("Argument 'det_spacing' is required with 'angles'.")
        vectors = angle_to_vector(det_spacing, angles)
    vectors = np.asarray(vectors, dtype=np.float64).reshape(-1, 12)
    pts = volume_coords_to_world_coords(np.concatenate([np.zeros((1, 3)), np.eye(3)]), input_shape)  # (4, 3)
    out = np.empty((len(vectors), 2, 4))
    for k, vec in enumerate(vectors):
        img = project_world_coordinates(pts, vec[0:3], vec[3:6], vec[6:9], vec[9:12], det_count)  # (4, 2)
        out[k, :, :3] = (img[1:] - img[0]).T
        out[k, :, 3] = img[0]
    # the pseudo-inverse leaves ~1e-17 where the geometry has exact zeros; they are far below half an
    # ulp of any fp32 coordinate they could be added to, and snapping them keeps axis-aligned
    # geometries (parallel3d) on the separable fast path
    lin = out[:, :, :3]
    lin[np.abs(lin) < 1e-13 * np.abs(lin).max(axis=(1, 2), keepdims=True)] = 0.0
    return out


def convert_from_scico_geometry(in_shape, matrices, det_shape) -> np.ndarray:
    """(V, 2, 4) projection matrices -> "parallel3d_vec" vectors (V, 12) (``_astra_3d.py:185-232``).
    As in the reference, the matrix rows themselves are taken as the detector steps v (rows) and u
    (columns), up to the axis-order flip: exact for orthonormal rows, i.e. unit detector spacing."""
    M = np.asarray(matrices, dtype=np.float64)
    lin, off = M[:, :, :3], M[:, :, 3]
    ray = np.cross(lin[:, 0], lin[:, 1])
    # detector centre: lift (detector centre index - image of the volume centre index) back to 3D
    vol_c = (np.asarray(in_shape, dtype=np.float64) - 1) / 2
    det_c = (np.asarray(det_shape, dtype=np.float64) - 1) / 2
    delta = det_c - (lin @ vol_c + off)                      # (V, 2)
    d = np.einsum("vmn,vm->vn", lin, delta)
    flip = lambda a: a[:, ::-1]  # noqa: E731  index (axis 0,1,2) order -> world (x, y, z) order
    return np.concatenate([flip(ray), flip(d), flip(lin[:, 1]), flip(lin[:, 0])], axis=1)
