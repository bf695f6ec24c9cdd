"""Pin the oracle against every known answer the reference's own tests hold for this path.

The reference (JAX) cannot run here, so these are the golden vectors / properties its test suite
asserts, restated against ``oracle/xray_np.py`` and the bit-identical C port ``oracle/xray_c.py``.
"""
import numpy as np
import pytest

from oracle import xray_c as C
from oracle import xray_np as O


def _setup_2d(nx, angles, dx=None, det_count=None):
    """Defaults of XRayTransform2D.__init__ (scico/linop/xray/_xray2d.py:88-120)."""
    if dx is None:
        dx = 2 * (np.sqrt(2) / 2,)
    if np.isscalar(dx):
        dx = 2 * (dx,)
    x0 = -(np.array(nx) * dx) / 2
    ny = int(np.ceil(np.linalg.norm(nx))) if det_count is None else det_count
    y0 = -ny / 2
    return O.view_table_2d(angles, x0, dx, y0), ny, dx


def _valid_adjoint(fwd, adj, in_shape, out_shape, seed=0):
    """scico/linop/_util.py:165-183 with white-noise vectors."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(in_shape).astype(np.float32)
    y = rng.standard_normal(out_shape).astype(np.float32)
    u, v = fwd(x).astype(np.float64), adj(y).astype(np.float64)
    a, b = float(np.sum(y * u)), float(np.sum(v * x))
    return abs(a - b) / max(abs(a), abs(b))


@pytest.mark.parametrize("impl", [O, C])
def test_3d_scaling_known_answers(impl):
    """scico/test/linop/xray/test_xray_3d.py:29-60 (assert_allclose, default rtol 1e-7)."""
    x = np.zeros((4, 4, 1), dtype=np.float32)
    x[1:3, 1:3, 0] = 1.0
    M = O.matrices_from_euler_angles(x.shape, (4, 4), "X", [[0.0]])
    truth = np.array([[[0, 0, 0, 0], [0, 1, 1, 0], [0, 1, 1, 0], [0, 0, 0, 0]]], dtype=np.float64)
    np.testing.assert_allclose(impl.project_3d(x, M, (4, 4)), truth)
    M = O.matrices_from_euler_angles(x.shape, (4, 4), "X", [[0.0]], voxel_spacing=[2.0, 1.0, 1.0])
    truth = np.array([[[0, 0.5, 0.5, 0]] * 4], dtype=np.float64)
    np.testing.assert_allclose(impl.project_3d(x, M, (4, 4)), truth)


@pytest.mark.parametrize("impl", [O, C])
def test_3d_matched_adjoint(impl):
    """scico/test/linop/xray/test_xray_3d.py:9-26: eps=1e-5, detector smaller than the volume."""
    N = 16
    det = int(N * 1.05 / np.sqrt(2.0))
    M = O.matrices_from_euler_angles((N,) * 3, (det, det), "X", np.linspace(0, np.pi, 3, endpoint=False)[:, None])
    err = _valid_adjoint(lambda x: impl.project_3d(x, M, (det, det)),
                         lambda y: impl.back_project_3d(y, M, (N,) * 3), (N,) * 3, (3, det, det))
    assert err < 1e-5


@pytest.mark.parametrize("impl", [O, C])
def test_2d_apply_adjoint(impl):
    """scico/test/linop/xray/test_xray_2d.py:33-74: shapes, default det_count, adjoint eps=1e-4."""
    nx = (12, 13)
    angles = np.linspace(0, np.pi, 10, endpoint=False)
    T, ny, _ = _setup_2d(nx, angles)
    assert ny == int(np.ceil(np.linalg.norm(nx))) == 18
    y = impl.project_2d(np.ones(nx, np.float32), T, ny)
    assert y.shape == (10, ny)
    # every pixel's two weights sum to 1 and nothing falls off the default detector
    np.testing.assert_allclose(y.sum(axis=1), 12 * 13, rtol=1e-6)
    err = _valid_adjoint(lambda x: impl.project_2d(x, T, ny), lambda s: impl.back_project_2d(s, T, nx), nx, (10, ny))
    assert err < 1e-4
    T, ny, _ = _setup_2d(nx, angles, det_count=14)
    assert impl.project_2d(np.ones(nx, np.float32), T, ny).shape == (10, 14)


@pytest.mark.parametrize("impl", [O, C])
def test_2d_matched_adjoint(impl):
    """scico/test/linop/xray/test_xray_2d.py:77-85 (issue #560): bins fall off the detector."""
    N = 16
    det_count = int(N * 1.05 / np.sqrt(2.0))
    angles = np.linspace(0, np.pi, 3, endpoint=False)
    T, ny, _ = _setup_2d((N, N), angles, dx=1.0 / np.sqrt(2), det_count=det_count)
    err = _valid_adjoint(lambda x: impl.project_2d(x, T, ny), lambda s: impl.back_project_2d(s, T, (N, N)), (N, N), (3, ny))
    assert err < 1e-5


def _psnr(ref, x):
    mse = np.mean((ref.astype(np.float64) - x) ** 2)
    return 10 * np.log10((ref.max() - ref.min()) ** 2 / mse)


@pytest.mark.parametrize("dx", [0.5, 1.0 / np.sqrt(2)])
@pytest.mark.parametrize("det_count_factor", [1.02 / np.sqrt(2.0), 1.0])
def test_2d_fbp(dx, det_count_factor):
    """scico/test/linop/xray/test_xray_2d.py:88-102: PSNR > 28 dB (C port projects, NumPy FBP math)."""
    N = 256
    x_gt = np.zeros((N, N), dtype=np.float32)
    x_gt[N // 4 : -N // 4, N // 4 : -N // 4] = 1.0
    det_count = int(det_count_factor * N)
    angles = np.linspace(0, np.pi, 360, endpoint=False)
    T, ny, dxx = _setup_2d((N, N), angles, dx=dx, det_count=det_count)
    y = C.project_2d(x_gt, T, ny)
    # FBP of _xray2d.py:158-197 with the C port as back projector
    V = len(angles)
    h = O.ramp_filter(ny)
    mask = C.back_project_2d(np.ones_like(y), T, (N, N)) >= np.float32(V * (1.0 - 1e-5))
    L = 2 * ny - 1
    hy = np.fft.ifft(np.fft.fft(h, n=L, axis=1) * np.fft.fft(y, n=L, axis=1), n=L, axis=1)
    hy = hy[:, (ny - 1) // 2 : -(ny - 1) // 2].real.astype(np.float32)
    x_fbp = np.float32(np.pi * dxx[0] * dxx[1] / V) * mask * C.back_project_2d(hy, T, (N, N))
    assert _psnr(x_gt, x_fbp) > 28


def test_2d_fbp_numpy_oracle_small():
    """Same FBP through the pure NumPy oracle at a size it finishes quickly."""
    N = 64
    x_gt = np.zeros((N, N), dtype=np.float32)
    x_gt[N // 4 : -N // 4, N // 4 : -N // 4] = 1.0
    angles = np.linspace(0, np.pi, 90, endpoint=False)
    T, ny, dx = _setup_2d((N, N), angles, dx=0.5, det_count=N)
    x_fbp = O.fbp_2d(C.project_2d(x_gt, T, ny), T, (N, N), dx)
    assert _psnr(x_gt, x_fbp) > 20


def test_geometry_known_answer():
    """scico/test/linop/xray/astra/test_astra_3d.py:200-222: the fixture geometry converts to
    [[[0,1,0,-2],[0,0,1,-1]]]; here: that matrix projects voxel (i,j,k) to pixel (j-2, k-1)."""
    M = np.array([[[0, 1, 0, -2], [0, 0, 1, -1]]], dtype=np.float32)
    x = np.zeros((3, 6, 5), dtype=np.float32)
    x[1, 4, 3] = 1.0
    p = O.project_3d(x, M, (4, 4))
    assert p[0, 2, 2] == 1.0 and p.sum() == 1.0


def test_euler_matches_scipy():
    """_xray3d.py:304 uses scipy Rotation.from_euler; the restatement must agree."""
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(1)
    for seq in ["X", "Y", "Z", "XY", "xy", "XYZ", "zyx", "ZX", "yz"]:
        ang = rng.uniform(-3, 3, (6, len(seq)))
        R = Rotation.from_euler(seq, ang).as_matrix()
        np.testing.assert_allclose(O._euler_to_matrices(seq, ang), R, atol=1e-14)


def test_quirk_exact_integer_left_edge():
    """_xray3d.py:224: ceil(left)-left is 0 when `left` is an integer, moving the whole weight
    to the NEXT bin.  Offsets of 0.25 make every left edge an exact integer."""
    M = np.array([[[1, 0, 0, -0.25], [0, 1, 0, -0.25]]], dtype=np.float32)
    x = np.zeros((3, 3, 1), dtype=np.float32)
    x[1, 1, 0] = 1.0
    p = O.project_3d(x, M, (4, 4))[0]  # left edge = 1.5 - .25 - .25 = 1.0 exactly on both axes
    assert p[2, 2] == 1.0 and p[1, 1] == 0.0
    np.testing.assert_array_equal(p, C.project_3d(x, M, (4, 4))[0])
