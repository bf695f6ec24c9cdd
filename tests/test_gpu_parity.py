"""Parity of the CUDA path (through the C ABI) against the oracle, on the B200.

Tolerances are BASELINE.json's: relative L2 <= 1e-5 in fp32 for both directions and
|<Ax,y> - <x,A^T y>| / (|Ax| |y|) < 1e-5.  Bin indices (int32) and the zero pattern of the
weights must be bit-exact; weights themselves agree to an ulp (asserted <= 2.4e-7 absolute).
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scico_b200 as sb
from scico_b200 import _lib
from scico_b200.xray import debug_weights_2d, debug_weights_3d
from oracle import xray_c as C
from oracle import xray_np as O

TOL = 1e-5
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(f for f in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(f).startswith("nb_"))  # nb_*: notebook tables


@pytest.fixture(scope="module")
def torch_dev(cuda_device):
    import torch

    return torch, cuda_device


def _gpu(torch, dev, op, x, adj=False):
    t = torch.as_tensor(np.ascontiguousarray(x), device=dev)
    return (op.adj(t) if adj else op(t)).cpu().numpy()


def _x_mats(N, D, V, **kw):
    return sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None], **kw)


CASES_3D = {
    "x16_det11": ((16, 16, 16), (11, 11), lambda: _x_mats((16,) * 3, (11, 11), 3)),
    "x_ragged": ((17, 18, 19), (20, 21), lambda: _x_mats((17, 18, 19), (20, 21), 5)),
    "x_wide": ((5, 70, 150), (5, 170), lambda: _x_mats((5, 70, 150), (5, 170), 11)),
    "x_tall": ((9, 150, 33), (12, 100), lambda: _x_mats((9, 150, 33), (12, 100), 8)),
    "x_single_view": ((4, 40, 40), (4, 57), lambda: _x_mats((4, 40, 40), (4, 57), 1)),
    "x_tiny": ((1, 1, 1), (1, 1), lambda: _x_mats((1, 1, 1), (1, 1), 2)),
    "x_unaligned_rows": ((32, 48, 40), (40, 70), lambda: _x_mats((32, 48, 40), (40, 70), 6, voxel_spacing=[0.8, 1, 1])),
    "x_det_offcentre": ((12, 30, 30), (20, 30), lambda: _x_mats((12, 30, 30), (20, 30), 4)),
    "xy_tilt": ((17, 18, 19), (20, 21), lambda: sb.matrices_from_euler_angles(
        (17, 18, 19), (20, 21), "XY", np.stack([np.linspace(0, np.pi, 5, endpoint=False), np.full(5, np.deg2rad(74.0))], 1))),
    "y_rot": ((20, 12, 24), (30, 14), lambda: sb.matrices_from_euler_angles(
        (20, 12, 24), (30, 14), "Y", np.linspace(0, np.pi, 5, endpoint=False)[:, None])),
    "along_axis0": ((3, 6, 5), (4, 4), lambda: np.array([[[0, 1, 0, -2], [0, 0, 1, -1]]], dtype=np.float64)),
    "magnified": ((8, 20, 20), (20, 50), lambda: _x_mats((8, 20, 20), (20, 50), 4, voxel_spacing=[2.0, 2.0, 2.0])),
}


# shapes inside the walk kernels' envelope (unit detector rows, D1 % 4 == 0)
CASES_3D.update({
    "walk_basic": ((20, 40, 48), (20, 64), lambda: _x_mats((20, 40, 48), (20, 64), 9)),
    "walk_ragged_tiles": ((13, 37, 45), (13, 68), lambda: _x_mats((13, 37, 45), (13, 68), 7)),
    "walk_small_det": ((9, 50, 50), (9, 32), lambda: _x_mats((9, 50, 50), (9, 32), 10)),
    "walk_det_rows_offcentre": ((12, 30, 30), (20, 44), lambda: _x_mats((12, 30, 30), (20, 44), 5)),
    "walk_one_view": ((8, 33, 31), (8, 48), lambda: _x_mats((8, 33, 31), (8, 48), 1)),
    "walk_two_views": ((8, 33, 31), (8, 48), lambda: _x_mats((8, 33, 31), (8, 48), 2)),
    # in-plane voxel spacing 1.4: minor-axis coefficient reaches 0.99 -> bins can jump by two (cold path)
    "walk_cold_spacing": ((6, 40, 44), (6, 96), lambda: _x_mats((6, 40, 44), (6, 96), 16, voxel_spacing=[1.0, 1.4, 1.4])),
    "walk_many_slices": ((70, 24, 20), (70, 36), lambda: _x_mats((70, 24, 20), (70, 36), 6)),
    # detector rows finer than the slices: every slice mixes two rows (joint forward, ROWS_MIX flush)
    "walk_row_mixing": ((12, 40, 44), (18, 96), lambda: _x_mats((12, 40, 44), (18, 96), 16, det_spacing=[0.75, 1.0])),
    "walk_row_mixing_full_turn": ((9, 70, 66), (8, 120),
                                  lambda: sb.matrices_from_euler_angles((9, 70, 66), (8, 120), "X", np.linspace(0, 2 * np.pi, 12, endpoint=False)[:, None],
                                                                        det_spacing=[1.3, 1.0])),
})


# general matrices (both detector coordinates depend on all three voxel indices): the brick kernels
def _tilt_mats(N, D, V, tilt_deg=74.0, seq="XY", **kw):
    ang = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, np.deg2rad(tilt_deg))], 1)
    return sb.matrices_from_euler_angles(N, D, seq, ang, **kw)


def _xyz_mats(N, D, V):
    rng = np.random.default_rng(5)
    return sb.matrices_from_euler_angles(N, D, "XYZ", rng.uniform(0, 2 * np.pi, size=(V, 3)))


CASES_3D.update({
    # several bricks per axis, ragged edges, detector rows 16-byte aligned (TMA-staged window, vector flush)
    "brick_tilt_aligned": ((21, 30, 37), (44, 48), lambda: _tilt_mats((21, 30, 37), (44, 48), 7)),
    # D1 % 4 != 0: cp.async-staged window, scalar flush
    "brick_tilt_odd_det": ((19, 22, 26), (33, 35), lambda: _tilt_mats((19, 22, 26), (33, 35), 6)),
    # random orientations: all three depth classes of the forward, rays along the volume diagonal included
    "brick_random_orientations": ((24, 20, 28), (40, 44), lambda: _xyz_mats((24, 20, 28), (40, 44), 24)),
    # detector smaller than the projected volume: zero fill outside the detector
    "brick_small_det": ((26, 26, 26), (16, 20), lambda: _tilt_mats((26, 26, 26), (16, 20), 5, tilt_deg=30.0)),
    # fine voxels (spacing 0.45): neighbouring lattice lanes share bins -> the forward's atomic variant
    "brick_fine_voxels": ((20, 20, 20), (16, 16), lambda: _tilt_mats((20, 20, 20), (16, 16), 6, voxel_spacing=[0.45, 0.45, 0.45])),
    # coarse voxels (spacing 2.6): the brick window does not fit -> thread-per-voxel kernels
    "brick_coarse_voxels_fallback": ((9, 10, 11), (40, 44), lambda: _tilt_mats((9, 10, 11), (40, 44), 4, voxel_spacing=[2.6, 2.6, 2.6])),
    # one brick, many views (views split over blockIdx.y)
    "brick_one_brick_many_views": ((8, 8, 8), (16, 16), lambda: _tilt_mats((8, 8, 8), (16, 16), 40)),
    # exact-integer left edges with a general matrix (the ceil quirk inside bins3)
    "brick_integer_edges": ((8, 8, 8), (12, 12), lambda: np.array(
        [[[1, 0, 0.5, 0.0], [0, 1, 0.5, 0.25]], [[0.5, 0.5, 0.5, 0.5], [0.25, -0.25, 1.0, 3.0]],
         [[0, 1, 1, -0.75], [1, 0, -1, 7.75]]], dtype=np.float64)),
})


def _quirk_mats():
    M = _x_mats((8, 12, 10), (9, 16), 3)
    M[:, :, 3] += 0.25
    return M


CASES_3D["quirk_integer_edges"] = ((8, 12, 10), (9, 16), _quirk_mats)


@pytest.mark.parametrize("force_general", [False, True], ids=["auto", "general"])
@pytest.mark.parametrize("name", list(CASES_3D))
def test_3d_parity(torch_dev, name, force_general):
    torch, dev = torch_dev
    N, D, mk = CASES_3D[name]
    M = mk()
    A = sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_FORCE_GENERAL if force_general else 0)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(N).astype(np.float32)
    y = rng.standard_normal(A.output_shape).astype(np.float32)
    Ax, ATy = _gpu(torch, dev, A, x), _gpu(torch, dev, A, y, adj=True)
    assert O.rel_l2(Ax, C.project_3d(x, A.matrices, D)) <= TOL
    assert O.rel_l2(ATy, C.back_project_3d(y, A.matrices, N)) <= TOL
    ns, ref = O.adjoint_gap(Ax, y, x, ATy)
    assert ns < TOL
    # index / weight arrays of _calc_weights, every view
    for v in range(len(M)):
        ul, w = debug_weights_3d(A, v)
        rul, rw = C.weights_3d(A.matrices[v], N, D)
        np.testing.assert_array_equal(w == 0, rw == 0)
        live = (rw != 0).any(axis=0)
        np.testing.assert_array_equal(ul[:, live], rul[:, live])
        assert np.abs(w - rw).max() <= 2.4e-7


def test_walk_kernels_selected_and_match_plane_kernels(torch_dev):
    """The walk kernels (xct_plane2.cuh) and the first-generation plane kernels are independent
    implementations of the same operator: bit-identical adjoint taps, same accumulation order."""
    torch, dev = torch_dev
    rng = np.random.default_rng(21)
    for name in ("walk_basic", "walk_ragged_tiles", "walk_small_det", "walk_det_rows_offcentre"):
        N, D, mk = CASES_3D[name]
        M = mk()
        A = sb.XRayTransform3D(N, M, D)
        B = sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_NO_WALK)
        assert A.plan_info()["adj_kernel"] == 2 and B.plan_info()["adj_kernel"] == 1
        assert A.plan_info()["fwd_kernel"] == 2 and B.plan_info()["fwd_kernel"] == 1
        y = rng.standard_normal(A.output_shape).astype(np.float32)
        a, b = _gpu(torch, dev, A, y, adj=True), _gpu(torch, dev, B, y, adj=True)
        assert O.rel_l2(a, b) <= 1e-6
        x = rng.standard_normal(N).astype(np.float32)
        assert O.rel_l2(_gpu(torch, dev, A, x), _gpu(torch, dev, B, x)) <= 2e-6
    # a sliced (unaligned) sinogram pointer falls back to the plane kernel and still agrees
    N, D, mk = CASES_3D["walk_basic"]
    A = sb.XRayTransform3D(N, mk(), D)
    y = rng.standard_normal(A.output_shape).astype(np.float32)
    buf = torch.zeros(y.size + 1, device=dev)
    buf[1:] = torch.as_tensor(y.ravel(), device=dev)
    assert buf[1:].data_ptr() % 16 != 0
    got = A.adj(buf[1:].view(A.output_shape)).cpu().numpy()
    assert O.rel_l2(got, C.back_project_3d(y, A.matrices, N)) <= TOL
    # ... and an unaligned OUTPUT pointer makes the forward fall back from the vector flush
    x = rng.standard_normal(N).astype(np.float32)
    obuf = torch.empty(int(np.prod(A.output_shape)) + 1, device=dev)
    out = obuf[1:].view(A.output_shape)
    assert out.data_ptr() % 16 != 0
    A.project(torch.as_tensor(x, device=dev), out=out)
    assert O.rel_l2(out.cpu().numpy(), C.project_3d(x, A.matrices, D)) <= TOL


def test_joint_forward_matches_single_column_walk_and_oracle(torch_dev):
    """walk_forward_joint_kernel (two columns per walk, three carried bins) against the one-column
    walk kernel and the oracle, over a full turn: every (major axis, minor sign, major sign) class,
    axis-aligned views (which stay on the one-column kernel) and ragged tiles."""
    torch, dev = torch_dev
    rng = np.random.default_rng(33)
    for N, D, angles in (
        ((12, 150, 140), (12, 216), np.linspace(0, 2 * np.pi, 24, endpoint=False)),
        ((5, 70, 131), (5, 192), np.linspace(0.01, 2 * np.pi + 0.01, 17, endpoint=False)),
        ((9, 64, 64), (9, 64), np.array([0.0, np.pi / 2, np.pi / 4, 3 * np.pi / 4, 1e-4, np.pi / 2 - 1e-4, 2.0])),
    ):
        M = sb.matrices_from_euler_angles(N, D, "X", np.asarray(angles)[:, None])
        A = sb.XRayTransform3D(N, M, D)
        B = sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_NO_JOINT)
        assert A.plan_info()["fwd_kernel"] == 2 and B.plan_info()["fwd_kernel"] == 2
        assert A.plan_info()["fwd_joint"] == 1 and B.plan_info()["fwd_joint"] == 0
        x = rng.standard_normal(N).astype(np.float32)
        a, b = _gpu(torch, dev, A, x), _gpu(torch, dev, B, x)
        ref = C.project_3d(x, A.matrices, D)
        assert O.rel_l2(a, b) <= 2e-6
        assert O.rel_l2(a, ref) <= TOL and O.rel_l2(b, ref) <= TOL
        # per view, so that a wrong class of a few views cannot hide in the norm
        for v in range(len(M)):
            assert O.rel_l2(a[v], ref[v]) <= TOL, v


def test_tile_forward_matches_register_stationary_joint_forward_and_oracle(torch_dev):
    """walk_forward_tile_kernel (CTA-shared 64 x 32 x 4 tile in shared memory, 8 views per CTA at a time) against
    walk_forward_joint_kernel (XCT_FLAG_NO_TILE) and the oracle: ragged tiles, a full turn (all eight view
    classes, E2 views next to the axes), detector rows off centre, a slab with slice / row offsets, few views
    (views split over blockIdx.y) and one view."""
    torch, dev = torch_dev
    rng = np.random.default_rng(31)
    cases = {name: CASES_3D[name] for name in ("walk_basic", "walk_ragged_tiles", "walk_small_det", "walk_det_rows_offcentre",
                                               "walk_one_view", "walk_many_slices")}
    cases["full_turn"] = ((10, 150, 170), (10, 232), lambda: sb.matrices_from_euler_angles(
        (10, 150, 170), (10, 232), "X", np.linspace(0, 2 * np.pi, 97, endpoint=False)[:, None]))
    for name, (N, D, mk) in cases.items():
        M = mk()
        A, B = sb.XRayTransform3D(N, M, D), sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_NO_TILE)
        assert A.analyse()["fwd_tile"] == 1 and B.analyse()["fwd_tile"] == 0 and B.plan_info()["fwd_joint"] == 1, name
        x = rng.standard_normal(N).astype(np.float32)
        ya, yb = _gpu(torch, dev, A, x), _gpu(torch, dev, B, x)
        assert O.rel_l2(ya, yb) <= 1e-6, name
        assert O.rel_l2(ya, C.project_3d(x, A.matrices, D)) <= TOL, name
        np.testing.assert_array_equal(ya == 0, yb == 0)
    # z-slab (what every rank of the slab-sharded operator runs)
    N, D, V = (96, 70, 90), (96, 128), 21
    M = _x_mats(N, D, V)
    S = sb.XRayTransform3D((40,) + N[1:], M, (40, D[1]), slice_offset=30, det_row_offset=30, det_rows_total=D[0])
    assert S.analyse()["fwd_tile"] == 1
    x = rng.standard_normal(N).astype(np.float32)
    want = C.project_3d(x[30:70], S.matrices, D, slice_offset=30)[:, 30:70]
    assert O.rel_l2(_gpu(torch, dev, S, x[30:70]), want) <= TOL


def test_tma_staged_adjoint_is_bit_identical_to_cp_async_staging(torch_dev):
    """The walk adjoint stages its sinogram window either with one TMA box per view (rows of
    consecutive slices consecutive; hardware zero fill at the detector edges) or with per-lane
    cp.async: same taps, same order, so the results must be bit-identical."""
    torch, dev = torch_dev
    rng = np.random.default_rng(5)
    for name in ("walk_basic", "walk_ragged_tiles", "walk_small_det", "walk_det_rows_offcentre", "walk_many_slices",
                 "walk_one_view", "walk_two_views"):
        N, D, mk = CASES_3D[name]
        M = mk()
        A = sb.XRayTransform3D(N, M, D)
        B = sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_NO_TMA)
        assert A.plan_info()["adj_tma"] == 1 and B.plan_info()["adj_tma"] == 0
        y = rng.standard_normal(A.output_shape).astype(np.float32)
        a, b = _gpu(torch, dev, A, y, adj=True), _gpu(torch, dev, B, y, adj=True)
        np.testing.assert_array_equal(a, b)
        assert O.rel_l2(a, C.back_project_3d(y, A.matrices, N)) <= TOL
    # z-slab plans (slice / detector-row offsets) go through the same box arithmetic
    N, D = (24, 40, 48), (24, 64)
    M = _x_mats(N, D, 9)
    y = rng.standard_normal((9,) + D).astype(np.float32)
    full = C.back_project_3d(y, M.astype(np.float32), N)
    for z0, z1 in ((0, 10), (10, 24), (3, 19)):
        kw = dict(slice_offset=z0, det_row_offset=z0, det_rows_total=D[0])
        A = sb.XRayTransform3D((z1 - z0,) + N[1:], M, (z1 - z0, D[1]), **kw)
        got = _gpu(torch, dev, A, np.ascontiguousarray(y[:, z0:z1]), adj=True)
        assert O.rel_l2(got, full[z0:z1]) <= TOL


def test_interleaved_adjoint_is_bit_identical_to_scalar_tap_adjoint(torch_dev):
    """walk_adjoint_vec_kernel (sinogram rows of four consecutive slices interleaved per bin by
    sino_interleave4_kernel, one LDS.128 per four slices and row step, 16 slices x 4 rows per thread)
    against walk_adjoint_kernel (XCT_FLAG_NO_ADJ_VEC) and the oracle: same taps in the same order, so the
    results must be bit-identical; ragged tiles, slice counts that are not multiples of 4 or 16, detector
    rows off centre (krow != 0, rows outside the detector), z-slab plans, repeated calls on new data
    (the interleaved scratch is reused), and slice sub-ranges through the pipelined host path."""
    torch, dev = torch_dev
    rng = np.random.default_rng(11)
    for name in ("walk_basic", "walk_ragged_tiles", "walk_small_det", "walk_det_rows_offcentre", "walk_many_slices",
                 "walk_one_view", "walk_two_views"):
        N, D, mk = CASES_3D[name]
        M = mk()
        A = sb.XRayTransform3D(N, M, D)
        B = sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_NO_ADJ_VEC)
        assert A.analyse()["adj_interleaved"] == 1 and B.analyse()["adj_interleaved"] == 0, name
        for _ in range(2):
            y = rng.standard_normal(A.output_shape).astype(np.float32)
            a, b = _gpu(torch, dev, A, y, adj=True), _gpu(torch, dev, B, y, adj=True)
            np.testing.assert_array_equal(a, b)
        assert O.rel_l2(a, C.back_project_3d(y, A.matrices, N)) <= TOL
        # host arrays: slice chunks of the pipelined path (sub-range launches at multiples of 4 slices)
        np.testing.assert_array_equal(np.asarray(A.T(y)), a)
    N, D = (24, 40, 48), (24, 64)
    M = _x_mats(N, D, 9)
    y = rng.standard_normal((9,) + D).astype(np.float32)
    full = C.back_project_3d(y, M.astype(np.float32), N)
    for z0, z1 in ((0, 10), (10, 24), (3, 19)):
        kw = dict(slice_offset=z0, det_row_offset=z0, det_rows_total=D[0])
        A = sb.XRayTransform3D((z1 - z0,) + N[1:], M, (z1 - z0, D[1]), **kw)
        assert A.analyse()["adj_interleaved"] == 1
        got = _gpu(torch, dev, A, np.ascontiguousarray(y[:, z0:z1]), adj=True)
        assert O.rel_l2(got, full[z0:z1]) <= TOL


def test_pair_is_cuda_graph_capturable(torch_dev):
    """xct_forward / xct_adjoint enqueue everything on the caller's stream without allocation or host
    synchronisation (the XLA FFI contract): a forward + adjoint pair captured into a CUDA
    graph replays on new data in the same buffers (3D walk kernels incl. the TMA box, and 2D)."""
    torch, dev = torch_dev
    rng = np.random.default_rng(3)
    N, D, mk = CASES_3D["walk_basic"]
    A3 = sb.XRayTransform3D(N, mk(), D)
    A2 = sb.XRayTransform2D((48, 40), np.linspace(0, np.pi, 12, endpoint=False))
    for A in (A3, A2):
        x = torch.as_tensor(rng.standard_normal(A.input_shape).astype(np.float32), device=dev)
        y = torch.empty(A.output_shape, device=dev)
        xb = torch.empty(A.input_shape, device=dev)
        A.project(x, out=y)
        A.back_project(y, out=xb)  # warm-up outside the capture (plans, kernel attributes)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            A.project(x, out=y)
            A.back_project(y, out=xb)
        x2 = rng.standard_normal(A.input_shape).astype(np.float32)
        x.copy_(torch.as_tensor(x2, device=dev))
        g.replay()
        torch.cuda.synchronize()
        want_y = A.project(torch.as_tensor(x2, device=dev))
        want_x = A.back_project(want_y)
        assert O.rel_l2(y.cpu().numpy(), want_y.cpu().numpy()) <= 1e-6
        assert O.rel_l2(xb.cpu().numpy(), want_x.cpu().numpy()) <= 1e-6


def test_2d_joint_forward_matches_plane_kernel_and_oracle(torch_dev):
    """walk2d_forward_joint_kernel (a lane walks two column pairs, three carried bins each) against the
    per-voxel plane kernel and the oracle: a full turn (all eight view classes), ragged tiles, an
    anisotropic pixel, a short detector, a batch; pixels wider than the envelope fall back."""
    torch, dev = torch_dev
    rng = np.random.default_rng(44)
    for nx, angles, kw in (
        ((150, 140), np.linspace(0, 2 * np.pi, 24, endpoint=False), {}),
        ((70, 131), np.linspace(0.01, 2 * np.pi + 0.01, 17, endpoint=False), dict(dx=(0.6, 0.7))),
        ((64, 64), np.array([0.0, np.pi / 2, np.pi / 4, 3 * np.pi / 4, 1e-4, np.pi / 2 - 1e-4, 2.0]), {}),
        ((96, 80), np.linspace(0, np.pi, 30, endpoint=False), dict(det_count=90)),
        ((33, 130), np.linspace(0, np.pi, 9, endpoint=False), dict(dx=0.95)),
    ):
        A = sb.XRayTransform2D(nx, angles, **kw)
        B = sb.XRayTransform2D(nx, angles, _flags=_lib.FLAG_NO_JOINT, **kw)
        assert A.plan_info()["fwd_joint"] == 1 and B.plan_info()["fwd_joint"] == 0
        x = rng.standard_normal(nx).astype(np.float32)
        a, b = _gpu(torch, dev, A, x), _gpu(torch, dev, B, x)
        ref = C.project_2d(x, A.view_table, A.ny)
        assert O.rel_l2(a, b) <= 2e-6
        for v in range(len(angles)):
            assert O.rel_l2(a[v], ref[v]) <= TOL, v
    A = sb.XRayTransform2D((40, 150), np.linspace(0, np.pi, 11, endpoint=False))
    xb = rng.standard_normal((3, 40, 150)).astype(np.float32)
    got = A.project(torch.as_tensor(xb, device=dev)).cpu().numpy()  # leading batch axis (the reference's vmap use)
    for k in range(3):
        assert O.rel_l2(got[k], C.project_2d(xb[k], A.view_table, A.ny)) <= TOL
    with pytest.warns(UserWarning):  # projected pixel wider than a bin: a coefficient can exceed 1
        W = sb.XRayTransform2D((48, 48), np.linspace(0, np.pi, 8, endpoint=False), dx=1.2)
    assert W.plan_info()["fwd_joint"] == 0
    x = rng.standard_normal((48, 48)).astype(np.float32)
    assert O.rel_l2(_gpu(torch, dev, W, x), C.project_2d(x, W.view_table, W.ny)) <= TOL


def test_2d_tile_forward_matches_register_stationary_forward_and_oracle(torch_dev):
    """walk2d_forward_tile_kernel (128 x 64 pixel tile in shared memory, one view per warp, 64-step walks of both
    column pairs) against walk2d_forward_joint_kernel (XCT_FLAG_NO_TILE) and the oracle.  Small problems normally take
    the one-launch kernel, so XCT_FLAG_2D_PER_CLASS selects the per-class launches (the large-problem path): a full
    turn (all eight view classes), ragged tiles in both axes, an anisotropic pixel, a short detector, one bin per
    pixel, odd widths (scalar tile loads), a batch."""
    torch, dev = torch_dev
    rng = np.random.default_rng(45)
    per_class = _lib.FLAG_2D_PER_CLASS
    for nx, angles, kw in (
        ((150, 140), np.linspace(0, 2 * np.pi, 24, endpoint=False), {}),
        ((70, 131), np.linspace(0.01, 2 * np.pi + 0.01, 17, endpoint=False), dict(dx=(0.6, 0.7))),
        ((64, 64), np.array([0.0, np.pi / 2, np.pi / 4, 3 * np.pi / 4, 1e-4, np.pi / 2 - 1e-4, 2.0]), {}),
        ((96, 80), np.linspace(0, np.pi, 30, endpoint=False), dict(det_count=90)),
        ((33, 130), np.linspace(0, np.pi, 9, endpoint=False), dict(dx=0.95)),
        ((260, 264), np.linspace(0, np.pi, 40, endpoint=False), {}),
    ):
        A = sb.XRayTransform2D(nx, angles, _flags=per_class, **kw)
        B = sb.XRayTransform2D(nx, angles, _flags=per_class | _lib.FLAG_NO_TILE, **kw)
        assert A.analyse()["fwd_tile"] == 1 and B.analyse()["fwd_tile"] == 0 and B.plan_info()["fwd_joint"] == 1
        x = rng.standard_normal(nx).astype(np.float32)
        a, b = _gpu(torch, dev, A, x), _gpu(torch, dev, B, x)
        ref = C.project_2d(x, A.view_table, A.ny)
        assert O.rel_l2(a, b) <= 2e-6
        for v in range(len(angles)):
            assert O.rel_l2(a[v], ref[v]) <= TOL, v
    A = sb.XRayTransform2D((40, 150), np.linspace(0, np.pi, 11, endpoint=False), _flags=per_class)
    xb = rng.standard_normal((3, 40, 150)).astype(np.float32)
    got = A.project(torch.as_tensor(xb, device=dev)).cpu().numpy()
    for k in range(3):
        assert O.rel_l2(got[k], C.project_2d(xb[k], A.view_table, A.ny)) <= TOL


def test_pipelined_host_path_with_short_edge_chunks(torch_dev, monkeypatch):
    """The host pipeline halves its first and last slice chunk (nothing overlaps the first H2D / the last D2H copy):
    with 32-slice chunks on 130 and 150 slices the ranges are 16, 32, ..., 18 and 16, 32, 32, 32, 16, 22; results
    equal the device entry points across every seam."""
    torch, dev = torch_dev
    monkeypatch.setenv("XCT_HOST_CHUNK_SLICES", "32")
    rng = np.random.default_rng(23)
    for N, D in (((130, 40, 48), (130, 72)), ((150, 36, 44), (140, 64))):
        M = _x_mats(N, D, 7)
        H = sb.XRayTransform3D(N, M, D)
        x = rng.standard_normal(N).astype(np.float32)
        y = rng.standard_normal(H.output_shape).astype(np.float32)
        fwd_dev, adj_dev = _gpu(torch, dev, H, x), _gpu(torch, dev, H, y, adj=True)
        out = np.full(H.output_shape, np.nan, np.float32)
        H.project(x, out=out)
        assert O.rel_l2(out, fwd_dev) <= 1e-6
        np.testing.assert_array_equal(out == 0, fwd_dev == 0)
        back = np.full(N, np.nan, np.float32)
        H.back_project(y, out=back)
        np.testing.assert_array_equal(back, adj_dev)


def test_brick_kernels_selected_and_match_thread_per_voxel_kernels(torch_dev):
    """General matrices run the brick kernels (TMA-staged adjoint window when detector rows are 16-byte aligned);
    they reproduce the thread-per-voxel family (XCT_FLAG_NO_BRICK) and the oracle, also through a z-slab with
    slice / detector-row offsets, and the expected forward classes are exercised."""
    torch, dev = torch_dev
    rng = np.random.default_rng(12)
    expect = {"brick_tilt_aligned": (3, 3, 1), "brick_tilt_odd_det": (3, 3, 0), "brick_random_orientations": (3, 3, 1),
              "brick_fine_voxels": (3, 3, 1), "brick_coarse_voxels_fallback": (0, 0, 0), "xy_tilt": (3, 3, 0)}
    for name, (fk, ak, tma) in expect.items():
        N, D, mk = CASES_3D[name]
        A = sb.XRayTransform3D(N, mk(), D)
        B = sb.XRayTransform3D(N, mk(), D, _flags=_lib.FLAG_NO_BRICK)
        ia, ib = A.plan_info(), B.plan_info()
        assert (ia["fwd_kernel"], ia["adj_kernel"], ia["adj_tma"]) == (fk, ak, tma), (name, ia)
        assert ib["fwd_kernel"] == 0 and ib["adj_kernel"] == 0 and ia["path_name"] == "3d_general"
        x = rng.standard_normal(N).astype(np.float32)
        y = rng.standard_normal(A.output_shape).astype(np.float32)
        assert O.rel_l2(_gpu(torch, dev, A, x), _gpu(torch, dev, B, x)) <= 1e-6, name
        assert O.rel_l2(_gpu(torch, dev, A, y, adj=True), _gpu(torch, dev, B, y, adj=True)) <= 1e-6, name
    # cp.async staging (XCT_FLAG_NO_TMA) equals the TMA staging bit for bit
    N, D, mk = CASES_3D["brick_tilt_aligned"]
    A, Cp = sb.XRayTransform3D(N, mk(), D), sb.XRayTransform3D(N, mk(), D, _flags=_lib.FLAG_NO_TMA)
    assert A.plan_info()["adj_tma"] == 1 and Cp.plan_info()["adj_tma"] == 0 and Cp.plan_info()["adj_kernel"] == 3
    y = rng.standard_normal(A.output_shape).astype(np.float32)
    np.testing.assert_array_equal(_gpu(torch, dev, A, y, adj=True), _gpu(torch, dev, Cp, y, adj=True))
    # z-slab of a tilted volume: slice offset + local detector rows (what the view-block sharding's per-slab plans use)
    M = mk()
    full_x = rng.standard_normal(N).astype(np.float32)
    full_y = rng.standard_normal((len(M),) + D).astype(np.float32)
    z0, z1, r0, r1 = 5, 16, 8, 40
    S = sb.XRayTransform3D((z1 - z0,) + N[1:], M, (r1 - r0, D[1]), slice_offset=z0, det_row_offset=r0, det_rows_total=D[0])
    assert S.plan_info()["adj_kernel"] == 3 and S.plan_info()["fwd_kernel"] == 3
    want_f = C.project_3d(full_x[z0:z1], S.matrices, D, slice_offset=z0)[:, r0:r1]
    assert O.rel_l2(_gpu(torch, dev, S, full_x[z0:z1]), want_f) <= TOL
    ypad = np.zeros_like(full_y)
    ypad[:, r0:r1] = full_y[:, r0:r1]
    want_a = C.back_project_3d(ypad, S.matrices, N)[z0:z1]
    assert O.rel_l2(_gpu(torch, dev, S, np.ascontiguousarray(full_y[:, r0:r1]), adj=True), want_a) <= TOL


def test_2d_one_launch_forward_matches_per_class_launches(torch_dev):
    """Small 2D problems run all eight view classes in ONE launch (walk2d_forward_joint_all_kernel); XCT_FLAG_2D_PER_CLASS
    keeps one launch per class.  Same results (up to the order of the REDs), full turn = all eight classes, batch."""
    torch, dev = torch_dev
    rng = np.random.default_rng(41)
    for nx, V, span, batch in (((512, 512), 360, np.pi, 1), ((96, 130), 64, 2 * np.pi, 1), ((130, 70), 40, 2 * np.pi, 3)):
        ang = np.linspace(0, span, V, endpoint=False)
        A, B = sb.XRayTransform2D(nx, ang), sb.XRayTransform2D(nx, ang, _flags=_lib.FLAG_2D_PER_CLASS)
        assert A.plan_info()["fwd_joint"] == 1 and B.plan_info()["fwd_joint"] == 1
        x = rng.standard_normal(((batch,) if batch > 1 else ()) + nx).astype(np.float32)
        L = _lib.lib()
        xt = torch.as_tensor(x, device=dev)
        L.xct_launch_count_reset()
        ya = A.project(xt).cpu().numpy()  # .project: accepts the leading batch axis
        la = L.xct_launch_count()
        L.xct_launch_count_reset()
        yb = B.project(xt).cpu().numpy()
        lb = L.xct_launch_count()
        assert la == 1 and lb >= 4, (la, lb)
        assert O.rel_l2(ya, yb) <= 1e-6
        x0 = x if batch == 1 else x[1]
        y0 = ya if batch == 1 else ya[1]
        assert O.rel_l2(y0, C.project_2d(x0, A.view_table, A.ny)) <= TOL


def test_3d_paths_selected(torch_dev):
    A = sb.XRayTransform3D((16,) * 3, _x_mats((16,) * 3, (16, 16), 4), (16, 16))
    assert A.plan_info()["path_name"] == "3d_sep" and A.plan_info()["row_aligned"] == 1
    N, D, mk = CASES_3D["xy_tilt"]
    assert sb.XRayTransform3D(N, mk(), D).plan_info()["path_name"] == "3d_general"


CASES_2D = {
    "12x13": dict(nx=(12, 13), V=10),
    "16x16_det11": dict(nx=(16, 16), V=3, dx=1.0 / np.sqrt(2), det_count=11),
    "64x64": dict(nx=(64, 64), V=90),
    "ragged": dict(nx=(101, 67), V=37, span=2 * np.pi),
    "dx_half": dict(nx=(300, 200), V=50, dx=0.5),
    "aniso": dict(nx=(90, 120), V=40, dx=(0.5, 0.7)),
    "single_pixel": dict(nx=(1, 1), V=4),
    "small_det": dict(nx=(80, 80), V=24, det_count=40),
    "big_det": dict(nx=(40, 40), V=24, det_count=200),
    "c2_512": dict(nx=(512, 512), V=360),
}


@pytest.mark.parametrize("force_general", [False, True], ids=["auto", "general"])
@pytest.mark.parametrize("name", list(CASES_2D))
def test_2d_parity(torch_dev, name, force_general):
    torch, dev = torch_dev
    kw = dict(CASES_2D[name])
    nx, V, span = kw.pop("nx"), kw.pop("V"), kw.pop("span", np.pi)
    angles = np.linspace(0, span, V, endpoint=False)
    A = sb.XRayTransform2D(nx, angles, _flags=_lib.FLAG_FORCE_GENERAL if force_general else 0, **kw)
    if name == "c2_512":
        assert A.ny == 725  # BASELINE.json configs[1]
    T = O.view_table_2d(angles, A.x0, A.dx, A.y0)
    np.testing.assert_array_equal(T, A.view_table)
    rng = np.random.default_rng(8)
    x = rng.standard_normal(nx).astype(np.float32)
    y = rng.standard_normal(A.output_shape).astype(np.float32)
    Ax, ATy = _gpu(torch, dev, A, x), _gpu(torch, dev, A, y, adj=True)
    assert O.rel_l2(Ax, C.project_2d(x, T, A.ny)) <= TOL
    assert O.rel_l2(ATy, C.back_project_2d(y, T, nx)) <= TOL
    ns, ref = O.adjoint_gap(Ax, y, x, ATy)
    assert ns < TOL
    for v in sorted({0, V // 3, V - 1}):
        inds, w = debug_weights_2d(A, v)
        rinds, rw = C.weights_2d(T[v], nx)
        np.testing.assert_array_equal(inds, rinds)
        assert np.abs(w - rw).max() <= 2.4e-7


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(f)[:-4] for f in GOLDEN])
def test_golden_fixtures(torch_dev, path):
    torch, dev = torch_dev
    g = np.load(path)
    if str(g["kind"]) == "2d":
        A = sb.XRayTransform2D(tuple(g["nx"]), g["angles"], dx=tuple(g["dx"]), det_count=int(g["det_count"]))
        if "table" in g.files:
            np.testing.assert_array_equal(A.view_table, g["table"])
        if "inds" in g.files:  # fixtures made by the reference's own source: its _calc_weights arrays
            for v in range(len(g["angles"])):
                inds, w = debug_weights_2d(A, v)
                np.testing.assert_array_equal(inds, g["inds"][v])
                assert np.abs(w - g["weights"][v]).max() <= 2.4e-7
    else:
        A = sb.XRayTransform3D(tuple(g["N"]), g["matrices"], tuple(g["D"]))
        ul, w = debug_weights_3d(A, len(g["matrices"]) // 2)
        live = (g["w_mid"] != 0).any(axis=0)
        np.testing.assert_array_equal(w == 0, g["w_mid"] == 0)
        np.testing.assert_array_equal(ul[:, live], g["ul_mid"][:, live])
        assert np.abs(w - g["w_mid"]).max() <= 2.4e-7
    assert O.rel_l2(_gpu(torch, dev, A, g["x"]), g["Ax"]) <= TOL
    assert O.rel_l2(_gpu(torch, dev, A, g["y"], adj=True), g["ATy"]) <= TOL
    if "fbp" in g.files:  # the reference source's own filtered back projection (_xray2d.py:158-197)
        got = A.fbp(torch.as_tensor(g["y"], device=dev)).cpu().numpy()
        assert O.rel_l2(got, g["fbp"]) <= 1e-4  # FFT round-off differs between libraries


def test_known_answers_exact(torch_dev):
    """scico/test/linop/xray/test_xray_3d.py:29-60 through the CUDA path (rtol 1e-7 as there)."""
    torch, dev = torch_dev
    x = np.zeros((4, 4, 1), np.float32)
    x[1:3, 1:3, 0] = 1.0
    H = sb.XRayTransform3D(x.shape, sb.matrices_from_euler_angles(x.shape, (4, 4), "X", [[0.0]]), (4, 4))
    np.testing.assert_allclose(_gpu(torch, dev, H, x), [[[0, 0, 0, 0], [0, 1, 1, 0], [0, 1, 1, 0], [0, 0, 0, 0]]])
    H = sb.XRayTransform3D(x.shape, sb.matrices_from_euler_angles(x.shape, (4, 4), "X", [[0.0]], voxel_spacing=[2.0, 1.0, 1.0]), (4, 4))
    np.testing.assert_allclose(_gpu(torch, dev, H, x), [[[0, 0.5, 0.5, 0]] * 4])


def test_reference_adjoint_tests(torch_dev):
    """valid_adjoint as the reference's tests call it (test_xray_2d.py:52-85, test_xray_3d.py:9-26),
    NumPy arrays in -> host-buffer C-ABI entry points."""
    A = sb.XRayTransform2D((12, 13), np.linspace(0, np.pi, 10, endpoint=False))
    assert sb.valid_adjoint(A, A.T, eps=1e-4)
    N = 16
    det = int(N * 1.05 / np.sqrt(2.0))
    A = sb.XRayTransform2D((N, N), np.linspace(0, np.pi, 3, endpoint=False), det_count=det, dx=1.0 / np.sqrt(2))
    assert sb.valid_adjoint(A, A.T, eps=1e-5)
    H = sb.XRayTransform3D((N,) * 3, _x_mats((N,) * 3, (det, det), 3), (det, det))
    assert sb.valid_adjoint(H, H.T, eps=1e-5)


def test_host_and_device_entry_points_agree(torch_dev):
    torch, dev = torch_dev
    rng = np.random.default_rng(9)
    H = sb.XRayTransform3D((6, 40, 50), _x_mats((6, 40, 50), (6, 70), 7), (6, 70))
    x = rng.standard_normal(H.input_shape).astype(np.float32)
    y = rng.standard_normal(H.output_shape).astype(np.float32)
    out = H(x)
    assert isinstance(out, np.ndarray) and out.dtype == np.float32
    assert O.rel_l2(out, _gpu(torch, dev, H, x)) <= 1e-6
    np.testing.assert_array_equal(H.adj(y), _gpu(torch, dev, H, y, adj=True))  # gather: deterministic
    # float64 input is accepted by __call__ (no dtype check) and rejected by adj
    assert O.rel_l2(H(x.astype(np.float64)), out) <= 1e-6
    with pytest.raises(ValueError):
        H.adj(y.astype(np.float64))


@pytest.mark.parametrize("N,D,spacing", [
    ((100, 40, 48), (100, 72), None),          # rows == slices
    ((96, 40, 48), (64, 72), None),            # detector shorter than the volume: slices fall off both ends
    ((70, 36, 44), (120, 64), None),           # detector taller: rows no slice touches stay zero
    ((80, 40, 48), (177, 72), [2.0, 1.0, 1.0]), # axis-0 scale 2: every second row only (offset keeps rows unit)
])
def test_pipelined_host_path_matches_device_path(torch_dev, N, D, spacing):
    """Host buffers of a 3D separable operator go through the chunked H2D / kernel / D2H pipeline
    (xct_*_host); the result must be what the device entry points give, chunk seams included."""
    torch, dev = torch_dev
    rng = np.random.default_rng(21)
    kw = {} if spacing is None else {"voxel_spacing": spacing}
    M = _x_mats(N, D, 9, **kw)
    H = sb.XRayTransform3D(N, M, D)
    H0 = sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_NO_HOST_PIPELINE)
    assert H.plan_info()["fwd_kernel"] == 2 and H.plan_info()["adj_kernel"] == 2  # walk kernels: pipeline eligible
    x = rng.standard_normal(N).astype(np.float32)
    y = rng.standard_normal(H.output_shape).astype(np.float32)
    fwd_dev, adj_dev = _gpu(torch, dev, H, x), _gpu(torch, dev, H, y, adj=True)
    assert O.rel_l2(fwd_dev, C.project_3d(x, H.matrices, D)) <= TOL
    for op in (H, H0):
        out = np.full(H.output_shape, np.nan, np.float32)  # stale contents must be overwritten everywhere
        op.project(x, out=out)
        assert O.rel_l2(out, fwd_dev) <= 1e-6
        np.testing.assert_array_equal(out == 0, fwd_dev == 0)
        back = np.full(N, np.nan, np.float32)
        op.back_project(y, out=back)
        np.testing.assert_array_equal(back, adj_dev)
    # page-locked buffers (the overlap case) give the same numbers
    xp = torch.empty(N, dtype=torch.float32, pin_memory=True)
    xp.copy_(torch.from_numpy(x))
    op_out = torch.empty(H.output_shape, dtype=torch.float32, pin_memory=True)
    H.project(xp.numpy(), out=op_out.numpy())
    assert O.rel_l2(op_out.numpy(), fwd_dev) <= 1e-6
    # a forward and an adjoint in flight together (xct_*_host_async), twice so that each direction's staging
    # buffers are reused while the other direction still runs; results only after host_wait()
    yp = torch.empty(H.output_shape, dtype=torch.float32, pin_memory=True)
    yp.copy_(torch.from_numpy(y))
    back_out = torch.empty(N, dtype=torch.float32, pin_memory=True)
    for op in (H, H0):
        for rep in range(2):
            op_out.fill_(float("nan"))
            back_out.fill_(float("nan"))
            assert op.project(xp.numpy(), out=op_out.numpy(), wait=False) is not None
            op.back_project(yp.numpy(), out=back_out.numpy(), wait=False)
            op.host_wait()
            assert O.rel_l2(op_out.numpy(), fwd_dev) <= 1e-6, rep
            np.testing.assert_array_equal(back_out.numpy(), adj_dev)
    with pytest.raises(ValueError):  # wait=False needs an explicit result buffer
        H.project(xp.numpy(), wait=False)


def test_forward_overwrites_output_and_is_linear(torch_dev):
    torch, dev = torch_dev
    rng = np.random.default_rng(10)
    A = sb.XRayTransform2D((70, 90), np.linspace(0, np.pi, 33, endpoint=False))
    x1 = torch.as_tensor(rng.standard_normal((70, 90)).astype(np.float32), device=dev)
    x2 = torch.as_tensor(rng.standard_normal((70, 90)).astype(np.float32), device=dev)
    y12 = A(2.0 * x1 - 3.0 * x2).cpu().numpy()
    lin = (2.0 * A(x1) - 3.0 * A(x2)).cpu().numpy()
    assert O.rel_l2(y12, lin) <= TOL
    assert float(A(torch.zeros_like(x1)).abs().max()) == 0.0


def test_2d_batch(torch_dev):
    torch, dev = torch_dev
    rng = np.random.default_rng(11)
    angles = np.linspace(0, np.pi, 21, endpoint=False)
    A = sb.XRayTransform2D((50, 60), angles)
    xb = rng.standard_normal((3, 50, 60)).astype(np.float32)
    yb = rng.standard_normal((3,) + A.output_shape).astype(np.float32)
    Ab = A.project(torch.as_tensor(xb, device=dev)).cpu().numpy()
    ATb = A.back_project(torch.as_tensor(yb, device=dev)).cpu().numpy()
    T = A.view_table
    for b in range(3):
        assert O.rel_l2(Ab[b], C.project_2d(xb[b], T, A.ny)) <= TOL
        assert O.rel_l2(ATb[b], C.back_project_2d(yb[b], T, (50, 60))) <= TOL
    with pytest.raises(ValueError):
        A(torch.as_tensor(xb, device=dev))  # __call__ enforces the exact input_shape


def test_2d_batch_adjoint_four_images_per_thread_is_bit_identical_to_single_images(torch_dev):
    """Batches of three and more images take plane_adjoint_kernel<Geom2, S = 4>: one coordinate / weight evaluation per
    (view, pixel) for four images.  Same taps in the same order per image, so the result equals image-by-image calls
    bit for bit; ragged batch (6 = 4 + 2), anisotropic pixels, a short detector (out-of-range bins)."""
    torch, dev = torch_dev
    rng = np.random.default_rng(21)
    for nx, V, kw, nb in (((70, 90), 33, {}, 6), ((130, 65), 48, dict(dx=(0.6, 0.45), det_count=61), 5), ((33, 200), 20, {}, 3)):
        A = sb.XRayTransform2D(nx, np.linspace(0, 2 * np.pi, V, endpoint=False), **kw)
        yb = torch.as_tensor(rng.standard_normal((nb,) + A.output_shape).astype(np.float32), device=dev)
        got = A.back_project(yb)
        for b in range(nb):
            one = A.back_project(yb[b].contiguous())
            d = (got[b] - one).double()
            assert float(torch.linalg.vector_norm(d) / torch.linalg.vector_norm(one.double())) <= 1e-6, (nx, b)
        assert O.rel_l2(got[nb - 1].cpu().numpy(), C.back_project_2d(yb[nb - 1].cpu().numpy(), A.view_table, nx)) <= TOL


def test_mass_conservation(torch_dev):
    """Per voxel the 2 (2D) / 4 (3D) weights sum to 1: each view's projection sums to sum(x)
    when nothing falls off the detector (SURVEY.md section 8a invariants)."""
    torch, dev = torch_dev
    A = sb.XRayTransform2D((12, 13), np.linspace(0, np.pi, 10, endpoint=False))
    y = _gpu(torch, dev, A, np.ones((12, 13), np.float32))
    np.testing.assert_allclose(y.sum(axis=1), 156.0, rtol=1e-5)
    N, D = (10, 30, 30), (10, 45)
    H = sb.XRayTransform3D(N, _x_mats(N, D, 9), D)
    x = np.random.default_rng(12).random(N).astype(np.float32)
    p = _gpu(torch, dev, H, x)
    np.testing.assert_allclose(p.sum(axis=(1, 2)), x.sum(dtype=np.float64), rtol=1e-5)


def test_fbp_psnr(torch_dev):
    """scico/test/linop/xray/test_xray_2d.py:88-102 on the CUDA path (device tensors, torch.fft)."""
    torch, dev = torch_dev
    N = 256
    x_gt = np.zeros((N, N), dtype=np.float32)
    x_gt[N // 4 : -N // 4, N // 4 : -N // 4] = 1.0
    for dx in (0.5, 1.0 / np.sqrt(2)):
        for f in (1.02 / np.sqrt(2.0), 1.0):
            A = sb.XRayTransform2D((N, N), np.linspace(0, np.pi, 360, endpoint=False), det_count=int(f * N), dx=dx)
            y = A(torch.as_tensor(x_gt, device=dev))
            x_fbp = A.fbp(y).cpu().numpy()
            mse = np.mean((x_gt.astype(np.float64) - x_fbp) ** 2)
            assert 10 * np.log10(1.0 / mse) > 28
    A = sb.XRayTransform2D((64, 64), np.linspace(0, np.pi, 90, endpoint=False), det_count=64)
    assert A.fbp(A(np.ones((64, 64), np.float32))).shape == (64, 64)  # NumPy path (test_fbp_jit shape)


def test_slab_decomposition_matches_full_operator(torch_dev):
    """z-slab sharding hook (_xray3d.py:212 slice_offset): slabs of the volume map to row blocks of
    the sinogram; the adjoint is bit-identical, the forward agrees to accumulation order."""
    torch, dev = torch_dev
    N, D, V = (24, 40, 36), (24, 52), 9
    M = _x_mats(N, D, V)
    H = sb.XRayTransform3D(N, M, D)
    rng = np.random.default_rng(13)
    x = rng.standard_normal(N).astype(np.float32)
    y = rng.standard_normal(H.output_shape).astype(np.float32)
    full_f, full_a = _gpu(torch, dev, H, x), _gpu(torch, dev, H, y, adj=True)
    for force in (0, _lib.FLAG_FORCE_GENERAL):
        parts_f, parts_a = [], []
        for z0, z1 in ((0, 8), (8, 16), (16, 24)):
            Hs = sb.XRayTransform3D((z1 - z0,) + N[1:], M, (z1 - z0, D[1]), slice_offset=z0, det_row_offset=z0,
                                    det_rows_total=D[0], _flags=force)
            parts_f.append(_gpu(torch, dev, Hs, x[z0:z1]))
            parts_a.append(_gpu(torch, dev, Hs, np.ascontiguousarray(y[:, z0:z1]), adj=True))
        assert O.rel_l2(np.concatenate(parts_f, axis=1), full_f) <= 1e-6
        slab_a = np.concatenate(parts_a, axis=0)
        if force == 0:
            np.testing.assert_array_equal(slab_a, full_a)
        else:
            assert O.rel_l2(slab_a, full_a) <= 1e-6


def test_view_block_decomposition(torch_dev):
    """View-block sharding: forward rows are independent per view; adjoint partials sum."""
    torch, dev = torch_dev
    angles = np.linspace(0, np.pi, 40, endpoint=False)
    A = sb.XRayTransform2D((96, 80), angles)
    rng = np.random.default_rng(14)
    x = rng.standard_normal((96, 80)).astype(np.float32)
    y = rng.standard_normal(A.output_shape).astype(np.float32)
    full_f, full_a = _gpu(torch, dev, A, x), _gpu(torch, dev, A, y, adj=True)
    acc = np.zeros((96, 80), np.float64)
    for v0, v1 in ((0, 13), (13, 26), (26, 40)):
        Ab = sb.XRayTransform2D((96, 80), angles[v0:v1], det_count=A.ny)
        assert O.rel_l2(_gpu(torch, dev, Ab, x), full_f[v0:v1]) <= 1e-6
        acc += _gpu(torch, dev, Ab, np.ascontiguousarray(y[v0:v1]), adj=True)
    assert O.rel_l2(acc, full_a) <= 1e-6
