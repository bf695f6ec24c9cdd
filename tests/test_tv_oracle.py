"""CPU checks of the TV / PDHG oracle (oracle/tv_np.py): the operator identities the reference's
own tests check for these pieces (scico/test/linop/test_diff.py adjoint identity,
scico/test/functional/test_norm.py prox optimality), and that PDHG built from them converges."""
import numpy as np

from oracle import tv_np as T
from oracle import xray_c as C
from oracle import xray_np as O


def test_finite_difference_matches_numpy_diff_and_adjoint_identity():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 6, 7)).astype(np.float32)
    d = T.finite_difference(x)
    assert d.shape == (3, 5, 6, 7)
    np.testing.assert_array_equal(d[0][:-1], np.diff(x, axis=0))
    np.testing.assert_array_equal(d[0][-1], 0)
    np.testing.assert_array_equal(d[2][..., :-1], np.diff(x, axis=2))
    z = rng.standard_normal(d.shape).astype(np.float32)
    a = np.sum(d.astype(np.float64) * z)
    b = np.sum(x.astype(np.float64) * T.finite_difference_adj(z))
    assert abs(a - b) / max(abs(a), abs(b)) < 1e-6


def test_l21_prox_is_the_minimiser_and_handles_zero():
    rng = np.random.default_rng(1)
    v = rng.standard_normal((3, 4, 5)).astype(np.float32)
    v[:, 0, 0] = 0
    lam = 0.7
    p = T.l21_prox(v, lam)
    assert np.all(p[:, 0, 0] == 0) and np.all(np.isfinite(p))

    def obj(u):
        return 0.5 * np.sum((u - v) ** 2) + lam * np.sum(np.sqrt((u ** 2).sum(axis=0)))

    for _ in range(20):
        assert obj(p) <= obj(p + 1e-2 * rng.standard_normal(p.shape).astype(np.float32)) + 1e-9
    # Moreau: prox_f(v) + prox_{f*}(v) = v with prox_{f*} the projection on the lam-ball per column
    cp = T.conj_prox(lambda w, l: T.l21_prox(w, lam * l), v, 1.0)
    nrm = np.sqrt((v ** 2).sum(axis=0, keepdims=True))
    np.testing.assert_allclose(cp, v * np.minimum(1.0, lam / np.maximum(nrm, 1e-30)), atol=1e-6)


def test_sql2_prox_closed_form():
    rng = np.random.default_rng(2)
    v, y = rng.standard_normal(9).astype(np.float32), rng.standard_normal(9).astype(np.float32)
    np.testing.assert_allclose(T.sql2_prox(v, y, 0.3), (0.3 * y + v) / 1.3, rtol=1e-6)
    cp = T.conj_prox(lambda w, l: T.sql2_prox(w, y, l), v, 0.5)
    np.testing.assert_allclose(cp, (v - 0.5 * y) / 1.5, rtol=1e-5, atol=1e-6)


def test_pdhg_oracle_decreases_objective_on_a_small_ct_problem():
    N, D, V = (6, 12, 12), (6, 18), 8
    M = O.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None]).astype(np.float32)
    A = lambda x: C.project_3d(x, M, D)
    AT = lambda y: C.back_project_3d(y, M, N)
    x_gt = np.zeros(N, np.float32)
    x_gt[2:4, 4:8, 3:9] = 1.0
    y = A(x_gt)
    x, z0, z1 = np.zeros(N, np.float32), np.zeros_like(y), np.zeros((3,) + N, np.float32)
    lam, tau, sigma = 0.05, 0.09, 0.09  # ||C||^2 = 95.7 here (power iteration): tau*sigma*||C||^2 = 0.78
    objs = [T.tv_objective(x, A, y, lam)]
    for _ in range(60):
        x, z0, z1 = T.pdhg_tv_step(x, z0, z1, A, AT, y, lam, tau, sigma)
        objs.append(T.tv_objective(x, A, y, lam))
    assert objs[-1] < 0.05 * objs[0]
    assert O.rel_l2(x, x_gt) < 0.1
