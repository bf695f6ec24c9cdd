"""Operator / LinearOperator protocol the projectors sit behind.

A thin restatement of the part of ``scico.operator._operator.Operator``
(``scico/operator/_operator.py:106-225``) and ``scico.linop._linop.LinearOperator``
(``scico/linop/_linop.py:126-440``) that the X-ray path and its callers use: same attribute
names, same shape / dtype checks and exceptions, ``.adj / .T / .H / .conj / .gram_op / @ ``
and scaling / sum algebra.  Arrays are NumPy arrays or torch tensors instead of jax arrays.
"""

from __future__ import annotations

from typing import Callable, Optional

import numpy as np


def _np_dtype(dt) -> np.dtype:
    """NumPy dtype of an array-like's dtype (accepts torch dtypes)."""
    try:
        return np.dtype(dt)
    except TypeError:
        name = str(dt).replace("torch.", "")
        return np.dtype(name)


def _shape_size(shape) -> int:
    return int(np.prod(shape)) if len(shape) else 1


class Operator:
    """Generic operator (``scico/operator/_operator.py:106``)."""

    def __init__(self, input_shape, output_shape=None, eval_fn: Optional[Callable] = None,
                 input_dtype=np.float32, output_dtype=None, jit: bool = False):
        self.input_shape = (input_shape,) if isinstance(input_shape, int) else tuple(input_shape)
        self.input_dtype = input_dtype
        if eval_fn:
            self._eval = eval_fn
        elif not hasattr(self, "_eval"):
            raise NotImplementedError(
                "Operator is an abstract base class when argument 'eval_fn' is not specified."
            )
        if output_shape is None or output_dtype is None:
            probe = self._eval(np.zeros(self.input_shape, dtype=input_dtype))
            if output_shape is None:
                output_shape = tuple(probe.shape)
            if output_dtype is None:
                output_dtype = _np_dtype(probe.dtype)
        self.output_shape = (output_shape,) if isinstance(output_shape, int) else tuple(output_shape)
        self.output_dtype = output_dtype
        self.input_size = _shape_size(self.input_shape)
        self.output_size = _shape_size(self.output_shape)
        self.shape = (self.output_shape, self.input_shape)
        self.matrix_shape = (self.output_size, self.input_size)
        if jit:
            self.jit()

    def jit(self):
        """No-op: kernels are precompiled (kept for API compatibility)."""

    def __repr__(self):
        return (f"{self.__module__}.{self.__class__.__qualname__}\n"
                f"  input_shape:  {self.input_shape}\n  output_shape: {self.output_shape}\n"
                f"  input_dtype:  {_np_dtype(self.input_dtype).name}\n"
                f"  output_dtype: {_np_dtype(self.output_dtype).name}\n")

    def __call__(self, x):
        if isinstance(x, Operator):
            if self.input_shape == x.output_shape:
                return Operator(input_shape=x.input_shape, output_shape=self.output_shape,
                                eval_fn=lambda z: self(x(z)), input_dtype=self.input_dtype,
                                output_dtype=x.output_dtype)
            raise ValueError(f"Incompatible shapes {self.shape}, {x.shape}.")
        if self.input_shape != tuple(x.shape):
            raise ValueError(
                f"Cannot evaluate {type(self)} with input_shape={self.input_shape} "
                f"on array with shape={tuple(x.shape)}."
            )
        return self._eval(x)


class LinearOperator(Operator):
    """Generic linear operator (``scico/linop/_linop.py:123``)."""

    def __init__(self, input_shape, output_shape=None, eval_fn: Optional[Callable] = None,
                 adj_fn: Optional[Callable] = None, input_dtype=np.float32, output_dtype=None,
                 jit: bool = False):
        super().__init__(input_shape=input_shape, output_shape=output_shape, eval_fn=eval_fn,
                         input_dtype=input_dtype, output_dtype=output_dtype, jit=False)
        if not hasattr(self, "_adj"):
            self._adj = None
        if not hasattr(self, "_gram"):
            self._gram = None
        if callable(adj_fn):
            self._adj = adj_fn
            self._gram = lambda x: self.adj(self(x))
        elif adj_fn is not None:
            raise TypeError(f"Argument 'adj_fn' must be either a Callable or None; got {adj_fn}.")

    # -- algebra (``_linop.py:198-262``) ------------------------------------------------
    def __add__(self, other):
        return self._binary(other, lambda a, b: a + b)

    def __sub__(self, other):
        return self._binary(other, lambda a, b: a - b)

    def _binary(self, other, op):
        if not isinstance(other, LinearOperator):
            raise TypeError(f"Operation not defined between {type(self)} and {type(other)}.")
        if self.shape != other.shape:
            raise ValueError(f"Shapes {self.shape} and {other.shape} do not match.")
        return LinearOperator(input_shape=self.input_shape, output_shape=self.output_shape,
                              eval_fn=lambda x: op(self(x), other(x)),
                              adj_fn=lambda y: op(self.adj(y), other.adj(y)),
                              input_dtype=self.input_dtype, output_dtype=self.output_dtype)

    def __mul__(self, other):
        if not np.isscalar(other):
            raise TypeError(f"Operation __mul__ not defined between {type(self)} and {type(other)}.")
        return LinearOperator(input_shape=self.input_shape, output_shape=self.output_shape,
                              eval_fn=lambda x: other * self(x),
                              adj_fn=lambda y: np.conj(other) * self.adj(y),
                              input_dtype=self.input_dtype, output_dtype=self.output_dtype)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if not np.isscalar(other):
            raise TypeError(f"Operation __truediv__ not defined between {type(self)} and {type(other)}.")
        return self * (1.0 / other)

    def __neg__(self):
        return self * (-1.0)

    def __matmul__(self, other):
        return self(other)

    def __rmatmul__(self, other):
        if isinstance(other, LinearOperator):
            return other(self)
        return self.adj(other.conj().T).conj().T

    def __call__(self, x):
        if isinstance(x, LinearOperator):
            if self.input_shape != x.output_shape:
                raise ValueError(f"Incompatible LinearOperator shapes {self.shape}, {x.shape}.")
            return LinearOperator(input_shape=x.input_shape, output_shape=self.output_shape,
                                  eval_fn=lambda z: self(x(z)), adj_fn=lambda z: x.adj(self.adj(z)),
                                  input_dtype=x.input_dtype, output_dtype=self.output_dtype)
        return super().__call__(x)

    def adj(self, y):
        """Adjoint applied to ``y``; dtype and shape are checked (``_linop.py:296-326``)."""
        if self._adj is None:
            raise NotImplementedError("adjoint not set: autodiff-derived adjoints need JAX")
        if isinstance(y, LinearOperator):
            return self.H(y)
        if _np_dtype(self.output_dtype) != _np_dtype(y.dtype):
            raise ValueError(f"Dtype error: expected {self.output_dtype}, got {y.dtype}.")
        if self.output_shape != tuple(y.shape):
            raise ValueError(
                f"Shapes do not conform: input array with shape {tuple(y.shape)} does not match "
                f"LinearOperator output_shape {self.output_shape}."
            )
        return self._adj(y)

    @property
    def T(self) -> "LinearOperator":
        return LinearOperator(input_shape=self.output_shape, output_shape=self.input_shape,
                              eval_fn=self.adj, adj_fn=self.__call__,
                              input_dtype=self.output_dtype, output_dtype=self.input_dtype)

    @property
    def H(self) -> "LinearOperator":
        return self.T  # real operators only on this path

    def conj(self) -> "LinearOperator":
        return LinearOperator(input_shape=self.input_shape, output_shape=self.output_shape,
                              eval_fn=lambda x: self(x.conj()).conj(),
                              adj_fn=lambda x: self.adj(x.conj()).conj(),
                              input_dtype=self.input_dtype, output_dtype=self.output_dtype)

    @property
    def gram_op(self) -> "LinearOperator":
        return LinearOperator(input_shape=self.input_shape, output_shape=self.input_shape,
                              eval_fn=self.gram, adj_fn=self.gram,
                              input_dtype=self.input_dtype, output_dtype=self.output_dtype)

    def gram(self, x):
        if self._gram is None:
            self._gram = lambda z: self.adj(self(z))
        return self._gram(x)


def valid_adjoint(A: LinearOperator, AT: LinearOperator, eps: Optional[float] = 1e-7, x=None, y=None,
                  key=None):
    """Adjoint test of ``scico/linop/_util.py:113-183`` with NumPy random vectors.

    Returns a bool (``eps`` given) or the relative error (``eps=None``)."""
    rng = np.random.default_rng(0 if key is None else key)
    if x is None:
        x = rng.standard_normal(A.input_shape).astype(_np_dtype(A.input_dtype))
    elif tuple(x.shape) != A.input_shape:
        raise ValueError("Shape of 'x' array not appropriate as an input for operator 'A'.")
    if y is None:
        y = rng.standard_normal(AT.input_shape).astype(_np_dtype(AT.input_dtype))
    elif tuple(y.shape) != AT.input_shape:
        raise ValueError("Shape of 'y' array not appropriate as an input for operator AT.")
    u = np.asarray(_to_numpy(A(x)), dtype=np.float64)
    v = np.asarray(_to_numpy(AT(y)), dtype=np.float64)
    yTu = float(np.sum(_to_numpy(y).astype(np.float64) * u))
    vTx = float(np.sum(v * _to_numpy(x).astype(np.float64)))
    err = abs(yTu - vTx) / max(abs(yTu), abs(vTx))
    return err if eps is None else err < eps


def _to_numpy(a):
    if isinstance(a, np.ndarray):
        return a
    return a.detach().cpu().numpy()


def power_iteration(A: LinearOperator, maxiter: int = 100, key=None):
    """Largest eigenvalue / eigenvector of a PSD operator (``scico/linop/_util.py:27-72``)."""
    rng = np.random.default_rng(0 if key is None else key)
    v = rng.standard_normal(A.input_shape).astype(np.float32)
    v = v / np.linalg.norm(v)
    mu = 0.0
    for _ in range(maxiter):
        Av = _to_numpy(A(v))
        mu = float(np.sum(v.astype(np.float64) * Av))
        v = (Av / np.linalg.norm(Av)).astype(np.float32)
    return mu, v


def operator_norm(A: LinearOperator, maxiter: int = 100, key=None) -> float:
    """``sqrt(lambda_max(A^H A))`` (``scico/linop/_util.py:75-110``)."""
    return float(np.sqrt(power_iteration(A.gram_op, maxiter=maxiter, key=key)[0]))
