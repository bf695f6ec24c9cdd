"""NumPy restatement of the TV / PDHG arithmetic around the projector pair (TEST INFRASTRUCTURE).

Restates, in float32 and in the reference's operation order:

* ``FiniteDifference(input_shape, append=0)``: ``scico/linop/_diff.py:25-96,236-272`` -- one
  ``snp.diff(x, axis, append=x[last])`` per axis, stacked to ``(ndim, *shape)``; its adjoint (the
  reference derives it by autodiff, ``scico/linop/_linop.py:187-192``) is written out by hand and
  checked against the forward with an inner-product test in ``tests/test_tv_oracle.py``.
* ``L21Norm.prox`` with ``l2_axis=0``: ``scico/functional/_norm.py:254-263`` (incl.
  ``no_nan_divide``, ``scico/numpy/util.py:302-315``).
* ``SquaredL2Loss(y).prox`` (identity forward operator): ``scico/loss.py:220-226``.
* ``Functional.conj_prox``: ``scico/functional/_functional.py:102-128``.
* ``PDHG.step`` for ``C = VerticalStack((A, D))``, ``g = Separable(SquaredL2Loss(y), lam*L21Norm)``,
  ``f = ZeroFunctional`` or ``NonNegativeIndicator``: ``scico/optimize/_primaldual.py:219-231``.

* ``scico.solver.cg``: ``scico/solver.py:367-405``.
* ``ADMM.step`` with ``LinearSubproblemSolver`` for ``f = SquaredL2Loss(y, A)``, ``g = lam*L21Norm``,
  ``C = FiniteDifference(append=0)`` (``examples/scripts/ct_tv_admm.py:68-84``):
  ``scico/optimize/_admm.py:334-378``, ``scico/optimize/_admmaux.py:206-269``.
* ``LinearizedADMM.step`` for ``C = VerticalStack((A, D))``: ``scico/optimize/_ladmm.py:253-277``.
* ``ProximalADMM.step`` for ``A = VerticalStack((C, alpha*D))``, ``B = -I``
  (``examples/scripts/ct_3d_tv_padmm.py:96-120``): ``scico/optimize/_padmm.py:349-363``.

Pinning: the reference's tests hold no golden vectors for these functions beyond generic operator
identities (``scico/test/linop/test_diff.py``, ``scico/test/functional/test_norm.py`` check the adjoint
identity and prox optimality); ``tests/test_tv_oracle.py`` checks the same identities here, the reference's
own ``cg`` / ``L21Norm.prox`` files executed over the NumPy ``jax`` stand-in give identical iterates, and
``padmm_tv_step`` together with the C projector pair reproduces the iteration-statistics table the reference
itself printed (real JAX / XLA) in ``data/notebooks/ct_3d_tv_padmm.ipynb``
(``tests/test_reference_notebook.py``, fixture ``tests/golden/nb_ct_3d_tv_padmm.npz``).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def finite_difference(x: np.ndarray) -> np.ndarray:
    """(ndim, *shape): d_a[i] = x[i+1] - x[i] along axis a, last entry 0 (append=0)."""
    x = np.asarray(x, dtype=f32)
    out = np.zeros((x.ndim,) + x.shape, dtype=f32)
    for a in range(x.ndim):
        last = np.take(x, [-1], axis=a)
        out[a] = np.diff(x, axis=a, append=last)
    return out


def finite_difference_adj(z: np.ndarray) -> np.ndarray:
    """Adjoint of :func:`finite_difference`: (D^T z)[i] = z_a[i-1] - z_a[i] with the last row of D
    zero (z_a[n-1] is multiplied by 0) and z_a[-1] = 0; summed over axes, left to right."""
    z = np.asarray(z, dtype=f32)
    nd = z.shape[0]
    out = np.zeros(z.shape[1:], dtype=f32)
    for a in range(nd):
        za = np.array(z[a], dtype=f32, copy=True)
        idx = [slice(None)] * nd
        idx[a] = -1
        za[tuple(idx)] = 0  # zero row of D
        lo = [slice(None)] * nd
        hi = [slice(None)] * nd
        lo[a] = slice(0, -1)
        hi[a] = slice(1, None)
        t = -za
        t[tuple(hi)] = t[tuple(hi)] + za[tuple(lo)]
        out = out + t
    return out


def l21_prox(v: np.ndarray, lam) -> np.ndarray:
    v = np.asarray(v, dtype=f32)
    lam = f32(lam)
    length = np.sqrt((np.abs(v) ** 2).sum(axis=0, keepdims=True)).astype(f32)
    direction = np.where(length != 0, v / np.where(length != 0, length, f32(1)), f32(0)).astype(f32)
    new_length = length - lam
    new_length = f32(0.5) * (new_length + np.abs(new_length))
    return (new_length * direction).astype(f32)


def sql2_prox(v: np.ndarray, y: np.ndarray, lam, scale=0.5) -> np.ndarray:
    c = f32(2.0 * scale * lam)
    return ((c * y + v) / (c + f32(1))).astype(f32)


def conj_prox(prox, v, lam):
    lam32 = f32(lam)
    return (v - lam32 * prox(v / lam32, 1.0 / lam)).astype(f32)


def pdhg_tv_step(x, z0, z1, A, AT, y, lam, tau, sigma, alpha=1.0, nonneg=False):
    """One PDHG iteration; returns (x, z0, z1).  A / AT: callables (projector pair)."""
    tau32, sig32, al32 = f32(tau), f32(sigma), f32(alpha)
    CTz = AT(z0) + finite_difference_adj(z1)  # VerticalStack adjoint: sum of block adjoints
    proxarg = x - tau32 * CTz
    x_new = np.maximum(proxarg, f32(0)) if nonneg else proxarg
    xbar = (f32(1) + al32) * x_new - al32 * x
    p0 = z0 + sig32 * A(xbar)
    p1 = z1 + sig32 * finite_difference(xbar)
    z0_new = conj_prox(lambda v, l: sql2_prox(v, y, l), p0, sigma)
    z1_new = conj_prox(lambda v, l: l21_prox(v, f32(lam) * f32(l)), p1, sigma)
    return x_new.astype(f32), z0_new, z1_new


def tv_objective(x, A, y, lam) -> float:
    r = (A(x) - y).astype(np.float64)
    d = finite_difference(x).astype(np.float64)
    return float(0.5 * np.sum(r * r) + lam * np.sum(np.sqrt((d * d).sum(axis=0))))


# ---------------------------------------------------------------------------------------------
# ADMM family (scico/optimize/_admm.py, _ladmm.py, _padmm.py) and CG (scico/solver.py)
# ---------------------------------------------------------------------------------------------
def cg(A, b, x0, tol=1e-5, atol=0.0, maxiter=1000):
    """``scico.solver.cg`` without preconditioner (``scico/solver.py:367-405``); float32 arrays,
    scalars as NumPy float32 like the reference's jax scalars.  Returns (x, info)."""
    b = np.asarray(b, dtype=f32)
    x = np.asarray(x0, dtype=f32)
    Ax = A(x)
    bn = np.linalg.norm(b.ravel()).astype(f32)
    r = (b - Ax).astype(f32)
    p = r
    num = np.sum(r * r, dtype=f32)
    ii = 0
    termination_tol_sq = np.maximum(f32(tol) * bn, f32(atol)) ** 2
    trace = [float(num)]  # <r, r> before every iteration (diagnostic, not part of the reference's info)
    while ii < maxiter and num > termination_tol_sq:
        Ap = A(p)
        alpha = f32(num / np.sum(p * Ap, dtype=f32))
        x = (x + alpha * p).astype(f32)
        r = (r - alpha * Ap).astype(f32)
        num_old = num
        num = np.sum(r * r, dtype=f32)
        beta = f32(num / num_old)
        p = (r + beta * p).astype(f32)
        ii += 1
        trace.append(float(num))
    return x, {"num_iter": ii, "rel_res": float(np.sqrt(num) / bn) if bn > 0 else 0.0, "trace": trace,
               "tol_sq": float(termination_tol_sq)}


def admm_tv_init(x0):
    """``ADMM.z_init`` / ``u_init`` (``_admm.py:297-332``): z = C x0, u = 0."""
    z = finite_difference(x0)
    return np.asarray(x0, dtype=f32), z, np.zeros_like(z)


def admm_tv_step(x, z, u, A, AT, y, lam, rho, cg_tol=1e-4, cg_maxiter=100):
    """One ADMM iteration for f = 1/2||Ax - y||^2, g = lam||.||_{2,1}, C = D, alpha = 1.
    Returns (x, z, u, cg_info)."""
    rho32 = f32(rho)
    # LinearSubproblemSolver.compute_rhs (_admmaux.py:231-255): 2*scale = 1, W = I
    rhs = np.zeros(x.shape, dtype=f32)
    rhs = rhs + f32(1.0) * AT(np.asarray(y, dtype=f32))
    rhs = (rhs + rho32 * finite_difference_adj(z - u)).astype(f32)

    def lhs(v):  # rho * C.gram_op + f.hessian  (_admmaux.py:218-227)
        return (rho32 * finite_difference_adj(finite_difference(v)) + AT(A(v))).astype(f32)

    x, info = cg(lhs, rhs, x, tol=cg_tol, maxiter=cg_maxiter)
    Cx = finite_difference(x)
    z_new = l21_prox(Cx + u, f32(lam) * f32(1.0 / rho))
    u_new = (u + Cx - z_new).astype(f32)
    return x, z_new, u_new, info


def ladmm_tv_init(x0, A):
    """``LinearizedADMM.z_init`` / ``u_init`` (``_ladmm.py:216-251``) for C = (A; D)."""
    x0 = np.asarray(x0, dtype=f32)
    z0, z1 = A(x0).astype(f32), finite_difference(x0)
    return x0, (z0, z1), (np.zeros_like(z0), np.zeros_like(z1))


def ladmm_tv_step(x, z, u, A, AT, y, lam, mu, nu, nonneg=False):
    """One linearized-ADMM iteration, C = VerticalStack((A, D)), g = Separable(SquaredL2Loss(y),
    lam*L21Norm), f = ZeroFunctional or NonNegativeIndicator.  z, u: (sinogram, gradient) pairs."""
    z0, z1 = z
    u0, u1 = u
    c = f32(mu / nu)
    t0 = A(x) - z0 + u0
    t1 = finite_difference(x) - z1 + u1
    proxarg = x - c * (AT(t0.astype(f32)) + finite_difference_adj(t1.astype(f32)))
    x_new = (np.maximum(proxarg, f32(0)) if nonneg else proxarg).astype(f32)
    Cx0, Cx1 = A(x_new).astype(f32), finite_difference(x_new)
    z0n = sql2_prox(Cx0 + u0, y, nu)
    z1n = l21_prox(Cx1 + u1, f32(lam) * f32(nu))
    u0n = (u0 + Cx0 - z0n).astype(f32)
    u1n = (u1 + Cx1 - z1n).astype(f32)
    return x_new, (z0n, z1n), (u0n, u1n)


def padmm_tv_init(x_shape, y_shape):
    """``ProximalADMMBase.__init__`` defaults (``_padmm.py:106-134``): x, z, u, u_old all zero."""
    x = np.zeros(x_shape, dtype=f32)
    z = (np.zeros(y_shape, dtype=f32), np.zeros((len(x_shape),) + tuple(x_shape), dtype=f32))
    u = (np.zeros_like(z[0]), np.zeros_like(z[1]))
    return x, z, u, (u[0].copy(), u[1].copy())


def padmm_tv_step(x, z, u, u_old, A, AT, y, lam, alpha, rho, mu, nu, nonneg=False):
    """One proximal-ADMM iteration for A_stack = (A; alpha*D), B = -I, c = 0,
    g = Separable(SquaredL2Loss(y), (lam/alpha)*L21Norm) (``ct_3d_tv_padmm.py:96-120``)."""
    al = f32(alpha)
    inv_mu, inv_nu = f32(1.0 / mu), f32(1.0 / nu)
    q0 = f32(2.0) * u[0] - u_old[0]
    q1 = f32(2.0) * u[1] - u_old[1]
    proxarg = x - inv_mu * (AT(q0.astype(f32)) + al * finite_difference_adj(q1.astype(f32)))
    x_new = (np.maximum(proxarg, f32(0)) if nonneg else proxarg).astype(f32)
    Ax0, Ax1 = A(x_new).astype(f32), (al * finite_difference(x_new)).astype(f32)
    p0 = z[0] + inv_nu * ((Ax0 - z[0]) + u[0])
    p1 = z[1] + inv_nu * ((Ax1 - z[1]) + u[1])
    plam = 1.0 / (rho * nu)
    z0n = sql2_prox(p0.astype(f32), y, plam)
    z1n = l21_prox(p1.astype(f32), f32(lam / alpha) * f32(plam))
    u0n = ((u[0] + Ax0) - z0n).astype(f32)
    u1n = ((u[1] + Ax1) - z1n).astype(f32)
    return x_new, (z0n, z1n), (u0n, u1n), (u[0], u[1])
