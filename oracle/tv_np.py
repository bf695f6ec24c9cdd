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

Pinning: these functions have no golden vectors in the reference's tests beyond generic operator
identities (``scico/test/linop/test_diff.py``, ``scico/test/functional/test_norm.py`` check the adjoint
identity and prox optimality); ``tests/test_tv_oracle.py`` checks the same identities here.
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
