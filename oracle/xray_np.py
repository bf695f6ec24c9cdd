"""NumPy restatement of SCICO's native X-ray projectors (TEST INFRASTRUCTURE).

This file is the *oracle*: a CPU restatement of the algorithm in the reference's
``scico/linop/xray/_xray2d.py`` and ``_xray3d.py``.  It exists only so that tests,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` leg of ``bench.py`` can check the
CUDA path.  Nothing under ``scico_b200/`` may import it.

Pinning status.  JAX is not installable in this image, so the reference cannot run on
its own array library.  Three pins instead:

1. The reference's OWN SOURCE FILES (``_xray2d.py``, ``_xray3d.py``, loaded unmodified from
   /root/reference) are executed over a NumPy stand-in for the few ``jax`` entry points
   they use (``oracle/jax_standin.py``); their forward / adjoint results, index and weight
   arrays are committed as ``tests/golden/ref_*.npz`` and this oracle (and the C port)
   reproduces them BIT FOR BIT (``tests/test_golden_oracle.py``,
   ``tests/test_reference_source.py``).
2. Every known answer the reference's own tests hold for this path
   (``tests/test_oracle_pins.py``):

* ``scico/test/linop/xray/test_xray_3d.py:29-60``  two exact 4x4 projections,
* ``scico/test/linop/xray/test_xray_3d.py:9-26``   matched adjoint <= 1e-5,
* ``scico/test/linop/xray/test_xray_2d.py:52-85``  adjoint tests (1e-4 / 1e-5, OOB bins),
* ``scico/test/linop/xray/test_xray_2d.py:88-102`` FBP PSNR > 28 dB (4 cases),
* ``scico/test/linop/xray/astra/test_astra_3d.py:200-222`` geometry known answer.

3. Output of the REAL reference (JAX / XLA on a GPU): the iteration-statistics table its ProximalADMM
   printed in the executed notebook ``data/notebooks/ct_3d_tv_padmm.ipynb`` (objective and residuals of
   1000 iterations, 4 significant digits; ``tests/golden/nb_ct_3d_tv_padmm.npz``).  The C port of this
   oracle plus ``oracle/tv_np.py`` rebuild that example and print the same numbers
   (``tests/test_reference_notebook.py``) -- the 3D projector pair inside a full solve.

What is NOT pinned bit for bit (and cannot be here): what only XLA decides -- its fp32 ``cos``/``sin``,
FMA contraction and scatter order.  The oracle (like the stand-in) *defines* those as strict
IEEE fp32, no contraction, NumPy float32 ``cos``/``sin``.

Every fp32 operation below is written as a separate NumPy float32 op so that rounding
happens exactly where the reference's expression tree rounds.
"""

from __future__ import annotations

import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------------------
# 2D: scico/linop/xray/_xray2d.py
# --------------------------------------------------------------------------------------
def view_table_2d(angles, x0, dx, y0) -> np.ndarray:
    """Per-view scalars (Pxmin, Pdx0, Pdx1, width), float32 (V, 4).

    Follows ``_xray2d.py:326-347`` (``_calc_weights`` before the per-pixel part).
    """
    angles = np.asarray(angles, dtype=np.float64).astype(f32)  # traced as f32, x64 off
    x0 = np.asarray(x0, dtype=np.float64).astype(f32)
    dx = np.asarray(dx, dtype=np.float64).astype(f32)
    y0 = f32(y0)
    c = np.cos(angles).astype(f32)  # _xray2d.py:326
    s = np.sin(angles).astype(f32)
    Px0 = (x0[0] * c + x0[1] * s) - y0  # :327
    Pdx0 = dx[0] * c  # :328
    Pdx1 = dx[1] * s
    cands = np.stack([Px0, Px0 + Pdx0, Px0 + Pdx1, (Px0 + Pdx0) + Pdx1])  # :329
    Pxmin = cands.min(axis=0)
    d1 = np.abs(Pdx0 + Pdx1)  # :342-343
    d2 = np.abs(Pdx0 - Pdx1)
    w = np.maximum(d1, d2)
    f = np.minimum(d1, d2)
    width = (w + f) / f32(2)  # :347
    return np.stack([Pxmin, Pdx0, Pdx1, width], axis=1).astype(f32)


def calc_weights_2d(table: np.ndarray, nx) -> tuple[np.ndarray, np.ndarray]:
    """(inds int32 (V,N0,N1), weights f32 (V,N0,N1)); ``_xray2d.py:331-351``."""
    table = np.asarray(table, dtype=f32)
    Pxmin = table[:, 0].reshape(-1, 1, 1)
    Pdx0 = table[:, 1].reshape(-1, 1, 1)
    Pdx1 = table[:, 2].reshape(-1, 1, 1)
    width = table[:, 3].reshape(-1, 1, 1)
    i = np.arange(nx[0], dtype=np.int32).astype(f32).reshape(1, -1, 1)
    j = np.arange(nx[1], dtype=np.int32).astype(f32).reshape(1, 1, -1)
    Px = (Pxmin + Pdx0 * i) + Pdx1 * j  # :331-335
    inds = np.floor(Px).astype(np.int32)  # :338
    dist = f32(1) - (Px - inds.astype(f32))  # :348
    weights = np.minimum(dist, width) / width  # :349
    return inds, weights.astype(f32)


def project_2d(im, table, ny: int) -> np.ndarray:
    """Forward projection ``_xray2d.py:248-265``. ``im`` (N0,N1) -> (V, ny)."""
    im = np.asarray(im, dtype=f32)
    V = table.shape[0]
    y = np.zeros((V, ny), dtype=f32)
    for a in range(V):  # per view to bound memory; per-view results independent
        inds, w = calc_weights_2d(table[a : a + 1], im.shape)
        inds, w = inds[0], w[0]
        wv = np.where((inds >= 0) & (inds < ny), w, f32(0))
        np.add.at(y[a], np.clip(inds, 0, ny - 1).ravel(), (im * wv).ravel())
        wv = np.where((inds + 1 >= 0) & (inds + 1 < ny), f32(1) - w, f32(0))
        np.add.at(y[a], np.clip(inds + 1, 0, ny - 1).ravel(), (im * wv).ravel())
    return y


def back_project_2d(y, table, nx) -> np.ndarray:
    """Back projection ``_xray2d.py:292-306``. ``y`` (V, ny) -> (N0, N1).

    Two separate sums over views, added at the end, each accumulated in view order.
    """
    y = np.asarray(y, dtype=f32)
    V, ny = y.shape
    s0 = np.zeros(nx, dtype=f32)
    s1 = np.zeros(nx, dtype=f32)
    for a in range(V):
        inds, w = calc_weights_2d(table[a : a + 1], nx)
        inds, w = inds[0], w[0]
        wv = np.where((inds >= 0) & (inds < ny), w, f32(0))
        s0 = s0 + y[a][np.clip(inds, 0, ny - 1)] * wv
        wv = np.where((inds + 1 >= 0) & (inds + 1 < ny), f32(1) - w, f32(0))
        s1 = s1 + y[a][np.clip(inds + 1, 0, ny - 1)] * wv
    return (s0 + s1).astype(f32)


def ramp_filter(n: int) -> np.ndarray:
    """Kak & Slaney eq. 61 ramp filter, tau = 1; ``_xray2d.py:199-221,175-177``."""
    x = np.arange(n) - (n - 1) // 2
    out = np.where(
        x == 0, 1.0 / 4.0, np.where(x % 2, -1.0 / (x.astype(np.float64) ** 2 * np.pi**2 + (x == 0)), 0.0)
    )
    return out.astype(f32).reshape(1, -1)


def fbp_2d(y, table, nx, dx) -> np.ndarray:
    """Filtered back projection ``_xray2d.py:158-197``."""
    y = np.asarray(y, dtype=f32)
    V, N = y.shape
    h = ramp_filter(N)
    mask = back_project_2d(np.ones_like(y), table, nx) >= f32(V * (1.0 - 1e-5))
    L = 2 * N - 1
    hf = np.fft.fft(h.astype(np.complex64), n=L, axis=1)
    yf = np.fft.fft(y.astype(np.complex64), n=L, axis=1)
    hy = np.fft.ifft(hf * yf, n=L, axis=1)[:, (N - 1) // 2 : -(N - 1) // 2].real.astype(f32)
    scale = f32(np.pi * dx[0] * dx[1] / V)
    return (scale * mask * back_project_2d(hy, table, nx)).astype(f32)


# --------------------------------------------------------------------------------------
# 3D: scico/linop/xray/_xray3d.py
# --------------------------------------------------------------------------------------
def calc_weights_3d(input_shape, matrix, det_shape, slice_offset: int = 0):
    """``_xray3d.py:206-266``. Returns (ul_ind (2,...) int32, w_ul, w_ur, w_ll, w_lr)."""
    M = np.asarray(matrix, dtype=f32)
    n0, n1, n2 = input_shape
    x0 = (np.arange(n0, dtype=np.int32).astype(f32) + f32(0.5)).reshape(-1, 1, 1)
    if slice_offset:
        x0 = x0 + f32(slice_offset)  # :212  (exact in fp32 for |values| < 2^23)
    x1 = (np.arange(n1, dtype=np.int32).astype(f32) + f32(0.5)).reshape(1, -1, 1)
    x2 = (np.arange(n2, dtype=np.int32).astype(f32) + f32(0.5)).reshape(1, 1, -1)

    def px(r):
        return ((M[r, 0] * x0 + M[r, 1] * x1) + M[r, 2] * x2) + M[r, 3]  # :216-217

    w = f32(0.5)
    out_ind = []
    tn = []
    for r in range(2):
        left = px(r) - f32(0.25)  # :223
        tn.append(np.minimum(np.ceil(left) - left, w).astype(f32))  # :224
        out_ind.append(np.floor(left).astype(np.int32))  # :225
    four = f32(4.0)
    w_ul = (tn[0] * tn[1]) * four  # :227-230
    w_ur = ((w - tn[0]) * tn[1]) * four
    w_ll = (tn[0] * (w - tn[1])) * four
    w_lr = ((w - tn[0]) * (w - tn[1])) * four
    r0, c0 = out_ind
    d0, d1 = det_shape

    def ok(r, c):
        return (r >= 0) & (r < d0) & (c >= 0) & (c < d1)

    w_ul = np.where(ok(r0, c0), w_ul, f32(0))  # :233-264
    w_ur = np.where(ok(r0 + 1, c0), w_ur, f32(0))
    w_ll = np.where(ok(r0, c0 + 1), w_ll, f32(0))
    w_lr = np.where(ok(r0 + 1, c0 + 1), w_lr, f32(0))
    return np.stack([r0, c0]), w_ul, w_ur, w_ll, w_lr


def project_3d(im, matrices, det_shape, slice_offset: int = 0) -> np.ndarray:
    """``_xray3d.py:110-159``. ``im`` (N0,N1,N2) -> (V, D0, D1)."""
    im = np.asarray(im, dtype=f32)
    matrices = np.asarray(matrices, dtype=f32)
    d0, d1 = det_shape
    out = np.zeros((len(matrices), d0, d1), dtype=f32)
    for v, M in enumerate(matrices):
        (r0, c0), w_ul, w_ur, w_ll, w_lr = calc_weights_3d(im.shape, M, det_shape, slice_offset)
        flat = out[v].reshape(-1)
        for dr, dc, w in ((0, 0, w_ul), (1, 0, w_ur), (0, 1, w_ll), (1, 1, w_lr)):  # :155-158
            r = r0 + dr
            c = c0 + dc
            keep = (r >= 0) & (r < d0) & (c >= 0) & (c < d1)  # mode="drop"
            np.add.at(flat, (r * d1 + c)[keep], (w * im)[keep])
    return out


def back_project_3d(proj, matrices, input_shape, slice_offset: int = 0) -> np.ndarray:
    """``_xray3d.py:161-204``. ``proj`` (V, D0, D1) -> (N0,N1,N2); views in order."""
    proj = np.asarray(proj, dtype=f32)
    matrices = np.asarray(matrices, dtype=f32)
    d0, d1 = proj.shape[1:]
    vol = np.zeros(input_shape, dtype=f32)
    for v, M in enumerate(matrices):
        (r0, c0), w_ul, w_ur, w_ll, w_lr = calc_weights_3d(input_shape, M, (d0, d1), slice_offset)
        y = proj[v]
        for dr, dc, w in ((0, 0, w_ul), (1, 0, w_ur), (0, 1, w_ll), (1, 1, w_lr)):  # :200-203
            r = np.clip(r0 + dr, 0, d0 - 1)  # JAX clamps OOB gather indices; weight is 0 there
            c = np.clip(c0 + dc, 0, d1 - 1)
            vol = vol + y[r, c] * w
    return vol


def matrices_from_euler_angles(
    input_shape, output_shape, seq, angles, degrees=False, voxel_spacing=None, det_spacing=None
) -> np.ndarray:
    """``_xray3d.py:268-327`` restated with an explicit Euler composition.

    Returns float64 (V, 2, 4), as the reference does (cast to f32 by the operator ctor).
    """
    angles = np.atleast_2d(np.asarray(angles, dtype=np.float64))
    if degrees:
        angles = np.deg2rad(angles)
    if voxel_spacing is None:
        voxel_spacing = np.ones(3)
    if det_spacing is None:
        det_spacing = np.ones(2)
    try:  # the reference's own call (_xray3d.py:304) when scipy is importable
        from scipy.spatial.transform import Rotation

        R = np.asarray(Rotation.from_euler(seq, angles).as_matrix()).reshape(-1, 3, 3)
    except ImportError:
        R = _euler_to_matrices(seq, angles)
    M = np.einsum("vmn,nn->vmn", R[:, :2, :], np.diag(np.asarray(voxel_spacing, dtype=np.float64)))  # :308-314
    M = np.einsum("mm,vmn->vmn", np.diag(1 / np.asarray(det_spacing, dtype=np.float64)), M)
    x0 = np.array(input_shape) / 2
    t = -np.einsum("vmn,n->vm", M, x0) + np.array(output_shape) / 2
    return np.concatenate([M, t[..., None]], axis=2)


def _axis_rot(axis: str, a: np.ndarray) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    z, o = np.zeros_like(a), np.ones_like(a)
    if axis == "x":
        m = [[o, z, z], [z, c, -s], [z, s, c]]
    elif axis == "y":
        m = [[c, z, s], [z, o, z], [-s, z, c]]
    elif axis == "z":
        m = [[c, -s, z], [s, c, z], [z, z, o]]
    else:
        raise ValueError(f"bad axis {axis!r}")
    return np.stack([np.stack(row, axis=-1) for row in m], axis=-2)


def _euler_to_matrices(seq: str, angles: np.ndarray) -> np.ndarray:
    """scipy ``Rotation.from_euler`` semantics: upper case = intrinsic, lower = extrinsic."""
    if len(seq) != angles.shape[1] or not 1 <= len(seq) <= 3:
        raise ValueError("seq / angles mismatch")
    intrinsic = seq.isupper()
    if not (intrinsic or seq.islower()):
        raise ValueError("cannot mix intrinsic and extrinsic rotations")
    R = np.broadcast_to(np.eye(3), (angles.shape[0], 3, 3)).copy()
    for n, ax in enumerate(seq.lower()):
        Rn = _axis_rot(ax, angles[:, n])
        R = R @ Rn if intrinsic else Rn @ R
    return R


# --------------------------------------------------------------------------------------
# helpers shared by tests
# --------------------------------------------------------------------------------------
def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


def adjoint_gap(Ax, y, x, ATy) -> tuple[float, float]:
    """(north_star form |<Ax,y>-<x,ATy>|/(|Ax||y|), reference form ``linop/_util.py:176-183``)."""
    Ax, y, x, ATy = (np.asarray(t, dtype=np.float64).ravel() for t in (Ax, y, x, ATy))
    a = float(Ax @ y)
    b = float(x @ ATy)
    ns = abs(a - b) / (np.linalg.norm(Ax) * np.linalg.norm(y) + 1e-300)
    ref = abs(a - b) / max(abs(a), abs(b), 1e-300)
    return ns, ref
