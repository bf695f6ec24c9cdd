"""The committed golden fixtures must be reproduced bit-for-bit by both oracles (CPU only).

Two families: ``xray*.npz`` were made from the oracle itself (``make_golden.py``); ``ref*.npz`` were
made by the REFERENCE'S OWN SOURCE executed over a NumPy stand-in for jax
(``make_reference_golden.py``, ``oracle/jax_standin.py``) -- those pin the restatement."""
import glob
import os

import numpy as np
import pytest

from oracle import xray_c as C
from oracle import xray_np as O

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(f for f in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(f).startswith("nb_"))  # nb_*: notebook tables


def test_fixture_inventory():
    assert len(FILES) == 19
    assert sum(os.path.basename(f).startswith("ref") for f in FILES) == 11


def _table(g):
    if "table" in g.files:
        return g["table"]
    return O.view_table_2d(g["angles"], g["x0"], g["dx"], float(g["y0"]))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracles_reproduce_golden(path):
    g = np.load(path)
    if str(g["kind"]) == "2d":
        nx, ny, T = tuple(g["nx"]), int(g["det_count"]), _table(g)
        for impl in (O, C):
            np.testing.assert_array_equal(impl.project_2d(g["x"], T, ny), g["Ax"])
            np.testing.assert_array_equal(impl.back_project_2d(g["y"], T, nx), g["ATy"])
        if "inds" in g.files:  # the reference's own _calc_weights arrays (_xray2d.py:308-351)
            inds, weights = O.calc_weights_2d(T, nx)
            np.testing.assert_array_equal(inds, g["inds"])
            np.testing.assert_array_equal(weights, g["weights"])
            for v in range(len(T)):
                ci, cw = C.weights_2d(T[v], nx)
                np.testing.assert_array_equal(ci, g["inds"][v])
                np.testing.assert_array_equal(cw, g["weights"][v])
        if "fbp" in g.files:  # _xray2d.py:158-197 (FFT round-off differs between libraries: not bitwise)
            got = O.fbp_2d(g["y"], T, nx, tuple(g["dx"]))
            assert O.rel_l2(got, g["fbp"]) <= 1e-5
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
