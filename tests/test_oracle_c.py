"""The C port must be bit-identical to the NumPy restatement (same op order)."""
import numpy as np
import pytest

from oracle import xray_c as C
from oracle import xray_np as O


@pytest.mark.parametrize("seq,tilt", [("X", None), ("XY", 74.0), ("ZYX", 31.0)])
def test_3d_bit_identical(seq, tilt):
    rng = np.random.default_rng(3)
    N, D, V = (9, 14, 11), (13, 12), 4
    ang = np.linspace(0, np.pi, V, endpoint=False)[:, None]
    if len(seq) > 1:
        ang = np.concatenate([ang] + [np.full((V, 1), np.deg2rad(tilt))] * (len(seq) - 1), axis=1)
    M = O.matrices_from_euler_angles(N, D, seq, ang).astype(np.float32)
    x = rng.standard_normal(N).astype(np.float32)
    y = rng.standard_normal((V,) + D).astype(np.float32)
    for off in (0, 5):
        np.testing.assert_array_equal(O.project_3d(x, M, D, off), C.project_3d(x, M, D, off))
        np.testing.assert_array_equal(O.back_project_3d(y, M, N, off), C.back_project_3d(y, M, N, off))
        assert O.rel_l2(C.project_3d(x, M, D, off, fused=True), O.project_3d(x, M, D, off)) < 1e-6
    ul, w = C.weights_3d(M[1], N, D)
    ul2, *w2 = O.calc_weights_3d(N, M[1], D)
    np.testing.assert_array_equal(ul, ul2)
    for q in range(4):
        np.testing.assert_array_equal(w[q], w2[q])


def test_2d_bit_identical():
    rng = np.random.default_rng(4)
    nx = (23, 17)
    angles = np.linspace(0, 2 * np.pi, 13, endpoint=False)
    dx = (0.6, 0.7)
    T = O.view_table_2d(angles, -(np.array(nx) * dx) / 2, dx, -12.0)
    ny = 24
    x = rng.standard_normal(nx).astype(np.float32)
    y = rng.standard_normal((13, ny)).astype(np.float32)
    np.testing.assert_array_equal(O.project_2d(x, T, ny), C.project_2d(x, T, ny))
    np.testing.assert_array_equal(O.back_project_2d(y, T, nx), C.back_project_2d(y, T, nx))
    inds, w = O.calc_weights_2d(T[3:4], nx)
    inds2, w2 = C.weights_2d(T[3], nx)
    np.testing.assert_array_equal(inds[0], inds2)
    np.testing.assert_array_equal(w[0], w2)
