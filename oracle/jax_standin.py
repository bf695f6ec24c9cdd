"""Run the reference's OWN projector source over a NumPy stand-in for ``jax`` (TEST INFRASTRUCTURE).

JAX cannot be installed in this image, so ``import scico`` fails.  The two files that hold the hot
path, ``scico/linop/xray/_xray2d.py`` and ``_xray3d.py``, use only a handful of ``jax`` /
``jax.numpy`` entry points, all of them element-wise NumPy look-alikes plus ``.at[].add``,
``jit``, ``vmap``, ``lax.map`` and ``lax.scan``.  This module provides those entry points on top of
NumPy with JAX's x64-disabled dtype rules (every float result is float32, every integer result
int32, Python scalars are weakly typed, arguments are canonicalised at the ``jit`` boundary) and
loads the two reference files UNMODIFIED from ``/root/reference`` (nothing is copied into this
repository).  What comes out is the reference's own expression trees evaluated in IEEE fp32 in
the order Python evaluates them: it pins the oracle's restatement (association order, masks,
index arithmetic, the ``lax.scan`` accumulation order) against the real source.  It does not pin
what only XLA decides: fused multiply-adds, its ``cos`` / ``sin``, the order of scatter updates.

Only ``tests/golden/make_reference_golden.py`` and ``tests/test_reference_source.py`` use this
module, and only in the container that has ``/root/reference``; the vectors they produce are
committed under ``tests/golden/ref_*.npz``.

Semantics taken from JAX's documentation:
* scatter ``x.at[idx].add(v)``: negative indices wrap, out-of-bounds updates are dropped, updates
  are applied in index order (``np.add.at``);
* gather ``x[idx]`` with integer arrays: negative indices wrap, out-of-bounds indices are clamped;
* ``astype(int)`` is int32; ``jnp.arange`` / ``jnp.mgrid`` are int32.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def _canon_dtype(dt):
    dt = np.dtype(dt)
    if dt == np.float64:
        return np.dtype(np.float32)
    if dt == np.int64:
        return np.dtype(np.int32)
    if dt == np.complex128:
        return np.dtype(np.complex64)
    return dt


def _wrap(a):
    """Canonicalise a NumPy result to JAX's x64-disabled dtypes and view it as a JArr."""
    if isinstance(a, tuple):
        return tuple(_wrap(v) for v in a)
    if isinstance(a, (np.ndarray, np.generic)):
        a = np.asarray(a)
        return a.astype(_canon_dtype(a.dtype), copy=False).view(JArr)
    return a


def _raw(a):
    return a.view(np.ndarray) if isinstance(a, JArr) else a


class _AtIndex:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def add(self, values, mode=None):  # functional scatter-add, out-of-bounds updates dropped
        out = np.array(_raw(self.arr), copy=True)
        idx = self.idx if isinstance(self.idx, tuple) else (self.idx,)
        idx = tuple(_raw(i) for i in idx)
        values = np.asarray(_raw(values), dtype=out.dtype)
        if all(isinstance(i, (int, np.integer)) for i in idx):
            out[idx] = out[idx] + values
            return _wrap(out)
        arrays = np.broadcast_arrays(*[np.asarray(i) for i in idx])
        shape = arrays[0].shape
        vals = np.broadcast_to(values, shape + out.shape[len(arrays):])
        keep = np.ones(shape, dtype=bool)
        norm = []
        for ax, ia in enumerate(arrays):
            n = out.shape[ax]
            ia = np.where(ia < 0, ia + n, ia)
            keep &= (ia >= 0) & (ia < n)
            norm.append(ia)
        sel = tuple(ia[keep] for ia in norm)
        np.add.at(out, sel, vals[keep])
        return _wrap(out)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self.arr, idx)


class JArr(np.ndarray):
    """ndarray with JAX's dtype canonicalisation after every operation, ``.at`` and clamped gathers."""

    __array_priority__ = 100

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ins = tuple(_raw(i) for i in inputs)
        res = getattr(ufunc, method)(*ins, **kwargs)
        return _wrap(res)

    def __array_function__(self, func, types_, args, kwargs):
        def strip(o):
            if isinstance(o, JArr):
                return o.view(np.ndarray)
            if isinstance(o, (list, tuple)):
                return type(o)(strip(v) for v in o)
            return o

        return _wrap(func(*strip(args), **{k: strip(v) for k, v in kwargs.items()}))

    @property
    def at(self):
        return _At(self)

    def astype(self, dtype, *a, **k):
        if dtype is int:
            dtype = np.int32
        elif dtype is float:
            dtype = np.float32
        return np.ndarray.astype(self.view(np.ndarray), _canon_dtype(dtype), *a, **k).view(JArr)

    def __getitem__(self, idx):
        items = idx if isinstance(idx, tuple) else (idx,)
        if any(isinstance(i, np.ndarray) and i.dtype.kind in "iu" for i in items):
            fixed = []
            ax = 0
            for i in items:
                if isinstance(i, np.ndarray) and i.dtype.kind in "iu":
                    n = self.shape[ax]
                    ii = _raw(i)
                    ii = np.where(ii < 0, ii + n, ii)
                    fixed.append(np.clip(ii, 0, n - 1))  # out-of-bounds gathers are clamped
                else:
                    fixed.append(_raw(i))
                ax += 1
            return _wrap(np.ndarray.__getitem__(self.view(np.ndarray), tuple(fixed)))
        return _wrap(np.ndarray.__getitem__(self.view(np.ndarray), tuple(_raw(i) for i in items) if isinstance(idx, tuple) else _raw(idx)))

    def __len__(self):
        return self.shape[0]


def _asarray(x, dtype=None, device=None):
    a = np.asarray(_raw(x) if not isinstance(x, (list, tuple)) else [_raw(v) for v in x])
    if dtype is not None:
        a = a.astype(_canon_dtype(int if dtype is int else dtype) if dtype is not int else np.int32)
    return _wrap(a)


def _canon_arg(a):
    """What jax.jit does to a non-static argument: arrays / scalars become canonical-dtype arrays,
    containers are mapped."""
    if isinstance(a, (list, tuple)):
        return type(a)(_canon_arg(v) for v in a)
    if isinstance(a, (np.ndarray, np.generic)):
        return _wrap(np.asarray(a))
    if isinstance(a, (float, int)) and not isinstance(a, bool):
        return _wrap(np.asarray(a))  # committed to f32 / i32 at the jit boundary
    return a


def _jit(fun=None, static_argnames=None, **_):
    if fun is None:
        return lambda f: _jit(f, static_argnames=static_argnames)
    static = {static_argnames} if isinstance(static_argnames, str) else set(static_argnames or ())
    import inspect

    names = list(inspect.signature(fun).parameters)

    def wrapped(*args, **kwargs):
        args = [a if names[i] in static else _canon_arg(a) for i, a in enumerate(args)]
        kwargs = {k: (v if (k in static or k == "device") else _canon_arg(v)) for k, v in kwargs.items()}
        return fun(*args, **kwargs)

    wrapped.__wrapped__ = fun
    return wrapped


def _vmap(fun, in_axes=0, **_):
    import functools

    @functools.wraps(fun)  # keeps the signature visible to _jit's static_argnames lookup
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(len(a) for a, ax in zip(args, axes) if ax is not None)
        outs = []
        for i in range(n):
            call = [(_wrap(np.asarray(_raw(a))[i]) if ax is not None else a) for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        if isinstance(outs[0], tuple):
            return tuple(_wrap(np.stack([_raw(o[k]) for o in outs])) for k in range(len(outs[0])))
        return _wrap(np.stack([_raw(o) for o in outs]))

    return mapped


class _MGrid:
    def __getitem__(self, key):
        return _wrap(np.mgrid[key].astype(np.int32))


def _build_jax():
    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    lax = types.ModuleType("jax.lax")
    typing_ = types.ModuleType("jax.typing")
    src = types.ModuleType("jax._src")
    src_lib = types.ModuleType("jax._src.lib")
    src_sh = types.ModuleType("jax._src.sharding")
    xc = types.SimpleNamespace(Device=type("Device", (), {}))
    src_lib.xla_client = xc
    src_sh.Sharding = type("Sharding", (), {})
    typing_.ArrayLike = object

    def f1(npf):
        return lambda x, *a, **k: _wrap(npf(np.asarray(_raw(x)), *a, **k))

    def f2(npf):
        return lambda x, y: _wrap(npf(_prep(x, y)[0], _prep(x, y)[1]))

    def _prep(x, y):
        # weak Python scalars take the other operand's dtype
        xr, yr = _raw(x), _raw(y)
        if isinstance(xr, (int, float)) and isinstance(yr, np.ndarray):
            xr = np.asarray(xr, dtype=yr.dtype if (isinstance(xr, int) or yr.dtype.kind == "f") else np.float32)
        if isinstance(yr, (int, float)) and isinstance(xr, np.ndarray):
            yr = np.asarray(yr, dtype=xr.dtype if (isinstance(yr, int) or xr.dtype.kind == "f") else np.float32)
        return np.asarray(xr), np.asarray(yr)

    jnp.asarray = _asarray
    jnp.array = _asarray
    jnp.zeros = lambda shape, dtype=np.float32, device=None: _wrap(np.zeros(shape, dtype=_canon_dtype(dtype)))
    jnp.ones_like = lambda x: _wrap(np.ones_like(_raw(x)))
    jnp.arange = lambda *a, **k: _wrap(np.arange(*a, **k).astype(np.int32))
    jnp.mgrid = _MGrid()
    jnp.stack = lambda xs, axis=0: _wrap(np.stack([np.asarray(_raw(v)) for v in xs], axis=axis))
    jnp.cos, jnp.sin = f1(np.cos), f1(np.sin)
    jnp.floor, jnp.ceil, jnp.abs = f1(np.floor), f1(np.ceil), f1(np.abs)
    jnp.min, jnp.max, jnp.sum = f1(np.min), f1(np.max), f1(np.sum)
    jnp.minimum, jnp.maximum = f2(np.minimum), f2(np.maximum)

    def where(c, a, b):
        a2, b2 = _prep(a, b)
        return _wrap(np.where(np.asarray(_raw(c)).astype(bool), a2, b2))

    jnp.where = where
    jnp.pi = np.pi
    jnp.float32 = np.float32
    fft = types.SimpleNamespace(
        fft=lambda x, n=None, axis=-1: _wrap(np.fft.fft(np.asarray(_raw(x)), n=n, axis=axis)),
        ifft=lambda x, n=None, axis=-1: _wrap(np.fft.ifft(np.asarray(_raw(x)), n=n, axis=axis)),
    )
    jnp.fft = fft

    def lax_map(f, xs, batch_size=None):
        return _wrap(np.stack([_raw(f(_wrap(np.asarray(_raw(xs))[i]))) for i in range(len(xs))]))

    def lax_scan(f, init, xs):
        carry = init
        n = len(xs[0]) if isinstance(xs, tuple) else len(xs)
        for i in range(n):
            item = tuple(_wrap(np.asarray(_raw(x))[i]) for x in xs) if isinstance(xs, tuple) else _wrap(np.asarray(_raw(xs))[i])
            carry, _ = f(carry, item)
        return carry, None

    lax.map, lax.scan = lax_map, lax_scan
    jax.jit, jax.vmap, jax.lax, jax.numpy = _jit, _vmap, lax, jnp
    jax.Device = xc.Device
    jax._src = src
    src.lib, src.sharding = src_lib, src_sh
    return {"jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.typing": typing_, "jax._src": src,
            "jax._src.lib": src_lib, "jax._src.sharding": src_sh}


def _build_scico_stubs():
    mods = {}
    for name in ("scico", "scico.numpy", "scico.numpy.util", "scico.typing", "scico.linop", "scico.linop._linop",
                 "scico.linop.xray"):
        m = types.ModuleType(name)
        m.__path__ = []  # packages
        mods[name] = m
    mods["scico.numpy"].Array = object
    mods["scico.numpy"].concatenate = lambda xs, axis=0: _wrap(np.concatenate([np.asarray(_raw(v)) for v in xs], axis=axis))
    mods["scico.numpy.util"].is_scalar_equiv = lambda s: np.isscalar(s) or (hasattr(s, "ndim") and s.ndim == 0)
    mods["scico.typing"].Shape = tuple
    mods["scico.typing"].DType = object

    class LinearOperator:  # the attributes the projector classes rely on (scico/linop/_linop.py:126-185)
        def __init__(self, input_shape, output_shape=None, eval_fn=None, adj_fn=None, input_dtype=np.float32,
                     output_dtype=None, jit=False, **kw):
            self.input_shape, self.output_shape = tuple(input_shape), tuple(output_shape)
            self.input_dtype, self.output_dtype = input_dtype, output_dtype or input_dtype
            self._eval, self._adj = eval_fn, adj_fn

        def __call__(self, x):
            return self._eval(x)

        def adj(self, y):
            return self._adj(y)

    mods["scico.linop._linop"].LinearOperator = LinearOperator
    mods["scico"].numpy = mods["scico.numpy"]
    return mods


_LOADED = {}


def _snp_module(mods):
    """The slice of ``scico.numpy`` (a jax.numpy wrapper in the reference) that ``solver.cg`` and
    ``L21Norm`` call, with the stand-in's dtype rules."""
    snp = mods["scico.numpy"]
    j = _build_jax()["jax.numpy"]
    for name in ("zeros", "sum", "abs", "where", "maximum", "minimum", "asarray", "array"):
        setattr(snp, name, getattr(j, name))
    snp.sqrt = lambda x: _wrap(np.sqrt(np.asarray(_raw(x))))
    snp.divide = lambda x, y: _wrap(np.divide(np.asarray(_raw(x)), np.asarray(_raw(y))))
    snp.count_nonzero = lambda x, **k: _wrap(np.count_nonzero(np.asarray(_raw(x)), **k))
    snp.BlockArray = type("BlockArray", (), {})
    snp.Array = object
    lin = types.ModuleType("scico.numpy.linalg")
    # jnp.linalg.norm of a real array: sqrt(sum(x * x)) in the array's dtype
    lin.norm = lambda x, *a, **k: _wrap(np.sqrt(np.sum(np.square(np.asarray(_raw(x))), dtype=np.asarray(_raw(x)).dtype)))
    snp.linalg = lin
    mods["scico.numpy.linalg"] = lin
    util = mods["scico.numpy.util"]
    util.no_nan_divide = lambda x, y: snp.where(y != 0, snp.divide(x, snp.where(y != 0, y, 1)), 0)  # scico/numpy/util.py:315
    util.is_complex_dtype = lambda d: np.dtype(d).kind == "c"
    util.is_real_dtype = lambda d: np.dtype(d).kind == "f"
    util.is_nested = lambda x: isinstance(x, (list, tuple)) and any(isinstance(v, (list, tuple)) for v in x)
    return snp


def load_reference_solver_pieces():
    """Returns (cg, L21Norm): ``scico/solver.py::cg`` and ``scico/functional/_norm.py::L21Norm``,
    loaded unmodified from /root/reference and executed over the stand-in (the other names those two
    files import are inert placeholders)."""
    if "cg" in _LOADED:
        return _LOADED["cg"], _LOADED["L21Norm"]
    if not available():
        raise RuntimeError("the reference checkout is not present (this only runs in the build container)")
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split(".")[0] in ("jax", "scico")}
    fakes = {**_build_jax(), **_build_scico_stubs()}
    _snp_module(fakes)
    jsl = types.ModuleType("jax.scipy.linalg")
    jsp = types.ModuleType("jax.scipy")
    jsp.linalg = jsl
    fakes["jax.scipy"], fakes["jax.scipy.linalg"] = jsp, jsl
    fakes["jax"].scipy = jsp
    for n in ("CircularConvolve", "ComposedLinearOperator", "Diagonal", "MatrixOperator", "Sum"):
        setattr(fakes["scico.linop"], n, type(n, (), {}))
    fakes["scico.linop"].LinearOperator = fakes["scico.linop._linop"].LinearOperator
    metric = types.ModuleType("scico.metric")
    metric.rel_res = lambda *a, **k: None
    fakes["scico.metric"] = metric
    fakes["scico.typing"].BlockShape = tuple
    fakes["scico.typing"].Axes = object
    fn_pkg = types.ModuleType("scico.functional")
    fn_pkg.__path__ = []
    fn_base = types.ModuleType("scico.functional._functional")

    class Functional:  # scico/functional/_functional.py: only the flags the subclasses set
        has_eval = False
        has_prox = False

        def __init__(self):
            pass

    fn_base.Functional = Functional
    fakes["scico.functional"], fakes["scico.functional._functional"] = fn_pkg, fn_base
    fakes["scico"].numpy = fakes["scico.numpy"]
    sys.modules.update(fakes)
    try:
        out = {}
        for tag, rel, name in (("solver", "solver.py", "scico.solver"), ("norm", os.path.join("functional", "_norm.py"), "scico.functional._norm")):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, "scico", rel))
            mod = importlib.util.module_from_spec(spec)
            mod.__package__ = name.rsplit(".", 1)[0]
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            out[tag] = mod
        _LOADED["cg"], _LOADED["L21Norm"] = out["solver"].cg, out["norm"].L21Norm
    finally:
        for k in list(sys.modules):
            if k in fakes or k in ("scico.solver", "scico.functional._norm"):
                if saved.get(k) is not None:
                    sys.modules[k] = saved[k]
                else:
                    sys.modules.pop(k, None)
    return _LOADED["cg"], _LOADED["L21Norm"]


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE_ROOT, "scico", "linop", "xray", "_xray3d.py"))


def load_reference_projectors():
    """Returns (XRayTransform2D, XRayTransform3D): the reference's classes, executed over the stand-in."""
    if _LOADED:
        return _LOADED["2d"], _LOADED["3d"]
    if not available():
        raise RuntimeError("the reference checkout is not present (this only runs in the build container)")
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "jax" or k.startswith("jax.") or k == "scico" or k.startswith("scico.")}
    fakes = {**_build_jax(), **_build_scico_stubs()}
    sys.modules.update(fakes)
    try:
        out = {}
        for tag, fname in (("2d", "_xray2d.py"), ("3d", "_xray3d.py")):
            path = os.path.join(REFERENCE_ROOT, "scico", "linop", "xray", fname)
            name = "scico.linop.xray." + fname[:-3]
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            mod.__package__ = "scico.linop.xray"
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            out[tag] = mod
        _LOADED["2d"], _LOADED["3d"] = out["2d"].XRayTransform2D, out["3d"].XRayTransform3D
    finally:
        for k in list(sys.modules):
            if k in fakes or k.startswith("scico.linop.xray._xray"):
                if saved.get(k) is not None:
                    sys.modules[k] = saved[k]
                else:
                    sys.modules.pop(k, None)
    return _LOADED["2d"], _LOADED["3d"]
