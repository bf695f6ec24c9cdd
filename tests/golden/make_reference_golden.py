"""Golden vectors produced by the REFERENCE'S OWN SOURCE (tests/golden/ref_*.npz).

``scico/linop/xray/_xray2d.py`` and ``_xray3d.py`` are loaded unmodified from /root/reference and
executed over the NumPy stand-in for jax in ``oracle/jax_standin.py`` (JAX itself cannot be installed
in this image).  Runs only in the build container, where /root/reference exists:

    python tests/golden/make_reference_golden.py

The files use the layout of the oracle-made fixtures (``make_golden.py``) plus ``source = "reference"``,
so every test that walks ``tests/golden/*.npz`` (oracle bit-for-bit reproduction on CPU, CUDA parity
on the GPU) also runs against what the reference source computes.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import jax_standin as J  # noqa: E402

A = np.asarray


def case_2d(R2, name, nx, angles, seed, out_dir=HERE, fbp=False, **kw):
    op = R2(nx, angles, **kw)
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(nx).astype(np.float32)
    y = rng.standard_normal(op.output_shape).astype(np.float32)
    inds, weights = R2._calc_weights(op.x0, op.dx, op.nx, op.angles, op.y0)
    width_free = {}
    if fbp:
        width_free["fbp"] = A(op.fbp(y), dtype=np.float32)
    # per-view scalars in the layout of oracle.view_table_2d, recovered from the reference's own arrays:
    # Pxmin = Px[0, 0], Pdx0 = Px[1, 0] - Px[0, 0] is NOT exact, so the table itself is taken from the
    # oracle by the tests; what is stored here are the reference's full index / weight arrays
    np.savez_compressed(
        os.path.join(out_dir, name + ".npz"), kind="2d", source="reference", nx=np.array(nx),
        angles=np.asarray(angles, dtype=np.float64), dx=np.array(op.dx, dtype=np.float64), det_count=op.ny,
        x0=np.asarray(op.x0, dtype=np.float64), y0=np.float64(op.y0),
        x=x, y=y, Ax=A(op.project(x), dtype=np.float32), ATy=A(op.back_project(y), dtype=np.float32),
        inds=A(inds, dtype=np.int32), weights=A(weights, dtype=np.float32), **width_free)


def case_3d(R3, name, N, D, M, seed, out_dir=HERE, x=None):
    op = R3(N, M, D)
    rng = np.random.default_rng(seed)
    if x is None:
        x = rng.standard_normal(N).astype(np.float32)
    y = rng.standard_normal(op.output_shape).astype(np.float32)
    mid = A(op.matrices)[len(M) // 2]
    ul, *w = R3._calc_weights(N, op.matrices[len(M) // 2], D)
    np.savez_compressed(
        os.path.join(out_dir, name + ".npz"), kind="3d", source="reference", N=np.array(N), D=np.array(D),
        matrices=np.asarray(M, dtype=np.float64), x=x, y=y, Ax=A(op.project(x), dtype=np.float32),
        ATy=A(op.back_project(y), dtype=np.float32), ul_mid=A(ul, dtype=np.int32),
        w_mid=np.stack([A(v, dtype=np.float32) for v in w]), matrix_mid=mid)


def generate(out_dir=HERE):
    R2, R3 = J.load_reference_projectors()
    pi = np.pi
    case_2d(R2, "ref2d_12x13_v10", (12, 13), np.linspace(0, pi, 10, endpoint=False), 30, out_dir)
    # test_xray_2d.py:77-85 shape: detector shorter than the image diagonal (out-of-bounds bins)
    case_2d(R2, "ref2d_16x16_v3_det11", (16, 16), np.linspace(0, pi, 3, endpoint=False), 31, out_dir,
            dx=1.0 / np.sqrt(2), det_count=int(16 * 1.05 / np.sqrt(2.0)))
    case_2d(R2, "ref2d_40x36_v24_fbp", (40, 36), np.linspace(0, pi, 24, endpoint=False), 32, out_dir, fbp=True)
    case_2d(R2, "ref2d_24x24_dx06_full_turn", (24, 24), np.linspace(0, 2 * pi, 16, endpoint=False), 33, out_dir, dx=(0.6, 0.7))

    x = np.zeros((4, 4, 1), np.float32)
    x[1:3, 1:3, 0] = 1.0  # test_xray_3d.py:29-60
    case_3d(R3, "ref3d_kat_default", (4, 4, 1), (4, 4), R3.matrices_from_euler_angles((4, 4, 1), (4, 4), "X", [[0.0]]), 40, out_dir, x=x)
    case_3d(R3, "ref3d_kat_voxel2", (4, 4, 1), (4, 4),
            R3.matrices_from_euler_angles((4, 4, 1), (4, 4), "X", [[0.0]], voxel_spacing=[2.0, 1.0, 1.0]), 41, out_dir, x=x)
    N, D = (16, 16, 16), (16, 24)
    case_3d(R3, "ref3d_16_x_v6", N, D, R3.matrices_from_euler_angles(N, D, "X", np.linspace(0, pi, 6, endpoint=False)[:, None]), 42, out_dir)
    N, D = (17, 18, 19), (20, 21)  # ct_projector_comparison_3d.py:44-51 style tilt
    angs = np.stack([np.linspace(0, pi, 5, endpoint=False), np.full(5, np.deg2rad(74.0))], 1)
    case_3d(R3, "ref3d_17x18x19_xy_tilt", N, D, R3.matrices_from_euler_angles(N, D, "XY", angs), 43, out_dir)
    N, D = (8, 12, 10), (9, 16)  # left edges on exact integers (_xray3d.py:224 quirk)
    M = np.array(R3.matrices_from_euler_angles(N, D, "X", np.linspace(0, pi, 3, endpoint=False)[:, None]), dtype=np.float64)
    M[:, :, 3] += 0.25
    case_3d(R3, "ref3d_quirk_integer_edges", N, D, M, 44, out_dir)
    N, D = (10, 20, 24), (14, 40)  # anisotropic voxels and detector pixels
    case_3d(R3, "ref3d_spacing", N, D, R3.matrices_from_euler_angles(
        N, D, "X", np.linspace(0, pi, 7, endpoint=False)[:, None], voxel_spacing=[1.0, 0.9, 0.8], det_spacing=[0.75, 0.6]), 45, out_dir)
    # inside the walk / joint / TMA envelope of the CUDA path (unit rows, D1 % 4 == 0, > 64 columns)
    N, D = (12, 72, 80), (12, 120)
    case_3d(R3, "ref3d_walk_12x72x80_v16", N, D, R3.matrices_from_euler_angles(N, D, "X", np.linspace(0, 2 * pi, 16, endpoint=False)[:, None]), 46, out_dir)


if __name__ == "__main__":
    if not J.available():
        raise SystemExit("needs /root/reference (build container only)")
    generate()
    print(sorted(f for f in os.listdir(HERE) if f.startswith("ref")))
