"""The ``ref*.npz`` fixtures are what the reference's own source computes (CPU only).

In the build container (where /root/reference exists) the two reference files are executed again
over the NumPy stand-in for jax and every stored array must come out bit-for-bit; elsewhere the
test is skipped and the committed vectors stand.  Also checks the drop-in geometry helper against the
reference's ``matrices_from_euler_angles`` (``_xray3d.py:268-327``)."""
import glob
import os

import numpy as np
import pytest

from oracle import jax_standin as J

HERE = os.path.dirname(os.path.abspath(__file__))
needs_reference = pytest.mark.skipif(not J.available(), reason="/root/reference is only present in the build container")


@needs_reference
def test_committed_reference_goldens_regenerate_bit_for_bit(tmp_path):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(HERE, "golden", "make_reference_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.generate(str(tmp_path))
    made = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "ref*.npz")))
    kept = sorted(os.path.basename(f) for f in glob.glob(os.path.join(HERE, "golden", "ref*.npz")))
    assert made == kept and len(made) == 11
    for name in made:
        a, b = np.load(tmp_path / name), np.load(os.path.join(HERE, "golden", name))
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            np.testing.assert_array_equal(a[k], b[k], err_msg=f"{name}:{k}")


@needs_reference
def test_reference_known_answers_through_its_own_source():
    """scico/test/linop/xray/test_xray_3d.py:29-60 evaluated by the reference source over the stand-in."""
    _, R3 = J.load_reference_projectors()
    x = np.zeros((4, 4, 1), np.float32)
    x[1:3, 1:3, 0] = 1.0
    M = R3.matrices_from_euler_angles((4, 4, 1), (4, 4), "X", [[0.0]])
    np.testing.assert_allclose(np.asarray(R3((4, 4, 1), M, (4, 4)).project(x))[0], [[0, 0, 0, 0], [0, 1, 1, 0], [0, 1, 1, 0], [0, 0, 0, 0]])
    M = R3.matrices_from_euler_angles((4, 4, 1), (4, 4), "X", [[0.0]], voxel_spacing=[2.0, 1.0, 1.0])
    np.testing.assert_allclose(np.asarray(R3((4, 4, 1), M, (4, 4)).project(x))[0], [[0, 0.5, 0.5, 0]] * 4)


@needs_reference
def test_drop_in_geometry_equals_the_reference_helper():
    import scico_b200 as sb

    _, R3 = J.load_reference_projectors()
    rng = np.random.default_rng(5)
    for seq in ("X", "XY", "zyx"):
        ang = rng.uniform(-3, 3, (6, len(seq)))
        for kw in ({}, dict(voxel_spacing=[1.0, 0.9, 0.8], det_spacing=[0.75, 0.6])):
            ref = np.asarray(R3.matrices_from_euler_angles((9, 10, 11), (12, 13), seq, ang, **kw))
            got = sb.matrices_from_euler_angles((9, 10, 11), (12, 13), seq, ang, **kw)
            np.testing.assert_array_equal(got.astype(np.float32), ref)  # the reference returns float32 (x64 off)


@needs_reference
def test_tv_oracle_pieces_equal_the_reference_source():
    """``scico/solver.py::cg`` (:367-405) and ``L21Norm.prox`` (``functional/_norm.py:254-263``) executed
    from the reference's files over the stand-in against the TV oracle's restatements, bit for bit."""
    from oracle import tv_np as T

    cg, L21Norm = J.load_reference_solver_pieces()
    rng = np.random.default_rng(0)
    v = rng.standard_normal((3, 6, 7, 8)).astype(np.float32)
    v[:, 0, 0, 0] = 0  # a zero-length group: no_nan_divide
    for lam in (0.3, 1.5):
        np.testing.assert_array_equal(np.asarray(L21Norm(l2_axis=0).prox(J._wrap(v), lam)), T.l21_prox(v, lam))
    n = 40
    Q = rng.standard_normal((n, n)).astype(np.float32)
    Amat = (Q @ Q.T + n * np.eye(n, dtype=np.float32)).astype(np.float32)
    b, x0 = rng.standard_normal(n).astype(np.float32), np.zeros(n, np.float32)
    for tol, maxiter in ((1e-4, 25), (1e-7, 100), (1e-30, 7)):
        xr, info = cg(lambda x: J._wrap((Amat @ np.asarray(x)).astype(np.float32)), J._wrap(b), J._wrap(x0), tol=tol, maxiter=maxiter)
        xo, io = T.cg(lambda x: (Amat @ x).astype(np.float32), b, x0, tol=tol, maxiter=maxiter)
        assert info["num_iter"] == io["num_iter"]
        np.testing.assert_array_equal(np.asarray(xr), xo)
        assert abs(float(info["rel_res"]) - io["rel_res"]) <= 1e-6 * max(io["rel_res"], 1e-30)


@needs_reference
def test_example_phantom_and_notebook_numbers_come_from_the_reference():
    """The tangle phantom of the notebook-pin tests equals ``scico/examples.py::create_tangle_phantom`` (the function is
    taken out of the reference file with ``ast`` and executed as it stands: plain NumPy, no jax in it) bit for bit, and
    ``nb_ct_3d_tv_padmm.npz`` regenerates from the reference's notebook."""
    import ast

    import _ct3d_example as E

    path = "/root/reference/scico/examples.py"
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "create_tangle_phantom")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    for shape in ((128, 256, 64), (7, 5, 9)):
        np.testing.assert_array_equal(E.tangle_phantom(*shape), ns["create_tangle_phantom"](*shape))

    import importlib.util

    spec = importlib.util.spec_from_file_location("make_notebook_golden", os.path.join(HERE, "golden", "make_notebook_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    kept = np.load(os.path.join(HERE, "golden", "nb_ct_3d_tv_padmm.npz"))
    keep = {k: kept[k].copy() for k in ("objective", "prml_rsdl", "dual_rsdl", "snr_db", "mae")}
    mod.main()  # rewrites the fixture in place from the notebook
    again = np.load(os.path.join(HERE, "golden", "nb_ct_3d_tv_padmm.npz"))
    for k, v in keep.items():
        np.testing.assert_array_equal(again[k], v, err_msg=k)
