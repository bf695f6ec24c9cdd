"""The committed golden fixtures must be reproduced bit-for-bit by both oracles (CPU only)."""
import glob
import os

import numpy as np
import pytest

from oracle import xray_c as C
from oracle import xray_np as O

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def test_fixture_inventory():
    assert len(FILES) == 8


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracles_reproduce_golden(path):
    g = np.load(path)
    if str(g["kind"]) == "2d":
        nx, ny, T = tuple(g["nx"]), int(g["det_count"]), g["table"]
        for impl in (O, C):
            np.testing.assert_array_equal(impl.project_2d(g["x"], T, ny), g["Ax"])
            np.testing.assert_array_equal(impl.back_project_2d(g["y"], T, nx), g["ATy"])
    else:
        N, D, M = tuple(g["N"]), tuple(g["D"]), g["matrices"].astype(np.float32)
        for impl in (O, C):
            np.testing.assert_array_equal(impl.project_3d(g["x"], M, D), g["Ax"])
            np.testing.assert_array_equal(impl.back_project_3d(g["y"], M, N), g["ATy"])
        ul, w = C.weights_3d(M[len(M) // 2], N, D)
        np.testing.assert_array_equal(ul, g["ul_mid"])
        np.testing.assert_array_equal(w, g["w_mid"])


def test_kat_fixtures_hold_the_reference_truth():
    """scico/test/linop/xray/test_xray_3d.py:40-60 truth matrices are what the fixtures store."""
    g = np.load(os.path.join(HERE, "golden", "xray3d_kat_default.npz"))
    np.testing.assert_allclose(g["Ax"][0], [[0, 0, 0, 0], [0, 1, 1, 0], [0, 1, 1, 0], [0, 0, 0, 0]])
    g = np.load(os.path.join(HERE, "golden", "xray3d_kat_voxel2.npz"))
    np.testing.assert_allclose(g["Ax"][0], [[0, 0.5, 0.5, 0]] * 4)
