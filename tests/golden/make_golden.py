"""Generate the golden fixtures under tests/golden/ from the NumPy oracle (oracle/xray_np.py).

The reference (JAX) cannot be imported in this image, so the vectors come from the oracle after it
has been pinned against the reference's own known answers (tests/test_oracle_pins.py).  Re-run:
    python tests/golden/make_golden.py
Shapes follow SURVEY.md section 8c.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import xray_np as O  # noqa: E402


def case_2d(name, nx, angles, seed, dx=None, det_count=None):
    if dx is None:
        dx = 2 * (np.sqrt(2) / 2,)
    if np.isscalar(dx):
        dx = 2 * (dx,)
    x0 = -(np.array(nx) * dx) / 2
    ny = int(np.ceil(np.linalg.norm(nx))) if det_count is None else det_count
    y0 = -ny / 2
    T = O.view_table_2d(angles, x0, dx, y0)
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(nx).astype(np.float32)
    y = rng.standard_normal((len(angles), ny)).astype(np.float32)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), kind="2d", nx=np.array(nx), angles=np.asarray(angles, dtype=np.float64),
        dx=np.array(dx, dtype=np.float64), det_count=ny, table=T, x=x, y=y,
        Ax=O.project_2d(x, T, ny), ATy=O.back_project_2d(y, T, nx),
    )


def case_3d(name, N, D, M, seed, x=None):
    M32 = np.asarray(M, dtype=np.float32)
    rng = np.random.default_rng(seed)
    if x is None:
        x = rng.standard_normal(N).astype(np.float32)
    y = rng.standard_normal((len(M32),) + tuple(D)).astype(np.float32)
    ul, *w = O.calc_weights_3d(N, M32[len(M32) // 2], D)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), kind="3d", N=np.array(N), D=np.array(D), matrices=np.asarray(M, dtype=np.float64),
        x=x, y=y, Ax=O.project_3d(x, M32, D), ATy=O.back_project_3d(y, M32, N), ul_mid=ul, w_mid=np.stack(w),
    )


if __name__ == "__main__":
    case_2d("xray2d_12x13_v10", (12, 13), np.linspace(0, np.pi, 10, endpoint=False), 10)
    case_2d("xray2d_16x16_v3_det11", (16, 16), np.linspace(0, np.pi, 3, endpoint=False), 11,
            dx=1.0 / np.sqrt(2), det_count=int(16 * 1.05 / np.sqrt(2.0)))
    case_2d("xray2d_64x64_v90", (64, 64), np.linspace(0, np.pi, 90, endpoint=False), 12)

    x = np.zeros((4, 4, 1), np.float32)
    x[1:3, 1:3, 0] = 1.0
    case_3d("xray3d_kat_default", (4, 4, 1), (4, 4), O.matrices_from_euler_angles((4, 4, 1), (4, 4), "X", [[0.0]]), 20, x=x)
    case_3d("xray3d_kat_voxel2", (4, 4, 1), (4, 4),
            O.matrices_from_euler_angles((4, 4, 1), (4, 4), "X", [[0.0]], voxel_spacing=[2.0, 1.0, 1.0]), 21, x=x)
    N = 16
    det = int(N * 1.05 / np.sqrt(2.0))
    case_3d("xray3d_16_x_v3", (N,) * 3, (det, det),
            O.matrices_from_euler_angles((N,) * 3, (det, det), "X", np.linspace(0, np.pi, 3, endpoint=False)[:, None]), 22)
    ang = np.stack([np.linspace(0, np.pi, 5, endpoint=False), np.full(5, np.deg2rad(74.0))], axis=1)
    case_3d("xray3d_17x18x19_xy_tilt", (17, 18, 19), (20, 21), O.matrices_from_euler_angles((17, 18, 19), (20, 21), "XY", ang), 23)
    # offsets of 0.25 force exact-integer left edges (the ceil quirk of _xray3d.py:224)
    M = O.matrices_from_euler_angles((8, 12, 10), (9, 16), "X", np.array([[0.0], [np.pi / 2], [0.7]]))
    M[:, :, 3] += 0.25
    case_3d("xray3d_quirk_integer_edges", (8, 12, 10), (9, 16), M, 24)
    print(sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
