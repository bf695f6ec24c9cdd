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


# ---------------------------------------------------------------------------------------------
# ADMM family restatements (oracle/tv_np.py: cg, admm_tv_step, ladmm_tv_step, padmm_tv_step)
# ---------------------------------------------------------------------------------------------
def _small_ct(nonzero_bg=False):
    N, D, V = (6, 12, 12), (6, 18), 8
    M = O.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None]).astype(np.float32)
    A = lambda x: C.project_3d(x, M, D)
    AT = lambda y: C.back_project_3d(y, M, N)
    x_gt = np.zeros(N, np.float32)
    x_gt[2:4, 4:8, 3:9] = 1.0
    return N, A, AT, x_gt, A(x_gt)


def test_cg_oracle_solves_an_spd_system_like_the_reference_test():
    """scico/test/test_solver.py::test_cg_standard: random SPD system, residual below tol."""
    rng = np.random.default_rng(12345)
    n = 24
    B = rng.standard_normal((n, n))
    S = (B @ B.T + n * np.eye(n)).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    x, info = T.cg(lambda v: S @ v, b, np.zeros(n, np.float32), tol=1e-6, maxiter=200)
    assert info["num_iter"] < 200 and info["rel_res"] <= 1e-6
    np.testing.assert_allclose(x, np.linalg.solve(S.astype(np.float64), b), rtol=2e-4, atol=2e-5)
    # maxiter = 0 returns x0 untouched, rel_res = ||b - A x0|| / ||b||
    x0 = rng.standard_normal(n).astype(np.float32)
    x_same, info0 = T.cg(lambda v: S @ v, b, x0, maxiter=0)
    np.testing.assert_array_equal(x_same, x0)
    assert info0["num_iter"] == 0


def test_admm_family_oracles_reach_the_pdhg_minimum():
    N, A, AT, x_gt, y = _small_ct()
    lam = 0.05
    # reference point: PDHG run long
    x, z0, z1 = np.zeros(N, np.float32), np.zeros_like(y), np.zeros((3,) + N, np.float32)
    for _ in range(400):
        x, z0, z1 = T.pdhg_tv_step(x, z0, z1, A, AT, y, lam, 0.09, 0.09)
    best = T.tv_objective(x, A, y, lam)

    xa, za, ua = T.admm_tv_init(np.zeros(N, np.float32))
    for _ in range(60):
        xa, za, ua, info = T.admm_tv_step(xa, za, ua, A, AT, y, lam, rho=1.0, cg_tol=1e-4, cg_maxiter=25)
    assert info["num_iter"] <= 25
    assert T.tv_objective(xa, A, y, lam) <= 1.02 * best + 1e-3

    cn2 = 95.7  # ||(A; D)||^2 for this geometry (power iteration)
    xl, zl, ul = T.ladmm_tv_init(np.zeros(N, np.float32), A)
    nu = 1.0
    mu = nu / (1.05 * cn2)
    for _ in range(1500):
        xl, zl, ul = T.ladmm_tv_step(xl, zl, ul, A, AT, y, lam, mu, nu)
    assert T.tv_objective(xl, A, y, lam) <= 1.05 * best + 1e-3

    xp, zp, up, uo = T.padmm_tv_init(N, y.shape)
    for _ in range(1500):
        xp, zp, up, uo = T.padmm_tv_step(xp, zp, up, uo, A, AT, y, lam, alpha=1.0, rho=1.0, mu=1.05 * cn2, nu=1.05)
    assert T.tv_objective(xp, A, y, lam) <= 1.05 * best + 1e-3
    for rec in (xa, xl, xp):
        assert O.rel_l2(rec, x_gt) < 0.12


def test_padmm_alpha_scaling_is_consistent():
    """A = (C; alpha D) with g1 = (lam/alpha) L21 is the same problem for every alpha
    (ct_3d_tv_padmm.py:96-104): both settings approach the same minimiser."""
    N, A, AT, x_gt, y = _small_ct()
    lam = 0.05
    recs = []
    for alpha, mu in ((1.0, 1.05 * 95.7), (3.0, 1.05 * (88.0 + 9 * 12.0))):
        x, z, u, uo = T.padmm_tv_init(N, y.shape)
        for _ in range(2500):
            x, z, u, uo = T.padmm_tv_step(x, z, u, uo, A, AT, y, lam, alpha=alpha, rho=1.0, mu=mu, nu=1.05)
        recs.append(x)
    assert O.rel_l2(recs[0], recs[1]) < 2e-2
