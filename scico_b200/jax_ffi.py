"""JAX binding of the projector pair: XLA FFI custom calls + linear transpose rules.

Import-guarded: JAX is not installed in the build image, so the JAX half of this module has never been
executed there (``tests/test_jax_ffi.py`` runs the reference's own acceptance tests through it the day a wheel
is present).  It mirrors the in-tree precedent for plugging an external projector behind ``LinearOperator``
(``scico/linop/xray/astra/_astra_3d.py:498-511``: forward and adjoint each wrapped so that differentiating one
calls the other), with ``jax.custom_derivatives.linear_call`` supplying both the JVP and the transpose.

What does NOT depend on JAX is in the C ABI and is tested without it (``xct_op_*``,
``include/scico_b200_xray.h``): the custom calls carry an operator ID, the native library keeps a copy of the
geometry and one plan per device, so a jitted call placed on device k uses device k's tables, a batch axis
added by ``jax.vmap`` is folded into the kernels' batch count, and an operator that was released yields an
error instead of a dangling pointer.  :class:`RegisteredOperator` is that id with a Python lifetime: it is
captured by the closures of :func:`ffi_pair`, so the operator lives as long as any jitted function built from it.

Usage (inside scico, where JAX is present)::

    from scico_b200.jax_ffi import ffi_pair
    proj, bproj = ffi_pair(sb.XRayTransform3D(input_shape, matrices, det_shape))
    # XRayTransform3D.__init__:  eval_fn=proj, adj_fn=bproj   (jit / grad / vjp / linear_transpose / vmap work)
"""
from __future__ import annotations

import ctypes
import os
import weakref

import numpy as np

from . import _lib

try:  # pragma: no cover - JAX is absent in the build image
    import jax
    import jax.numpy as jnp
    from jax.custom_derivatives import linear_call

    HAVE_JAX = True
except Exception:  # pragma: no cover
    HAVE_JAX = False

_HERE = os.path.dirname(os.path.abspath(__file__))
FFI_LIB = os.path.join(_HERE, "libscico_b200_ffi.so")
_registered = False


class RegisteredOperator:
    """An operator of :mod:`scico_b200.xray` in the native registry (``xct_op_register_2d/3d``): an integer id
    that stays valid until the last Python reference is gone (``xct_op_release`` runs in a finalizer)."""

    def __init__(self, op):
        L = _lib.lib()
        geom = op._geom(0)  # the device field is ignored by the registry: plans are per device, made on demand
        oid = ctypes.c_int64()
        fn = L.xct_op_register_3d if len(op.input_shape) == 3 else L.xct_op_register_2d
        _lib.check(fn(ctypes.byref(geom), ctypes.byref(oid)))
        self.id = int(oid.value)
        self.input_shape, self.output_shape = tuple(op.input_shape), tuple(op.output_shape)
        self._finalizer = weakref.finalize(self, L.xct_op_release, ctypes.c_int64(self.id))

    def plan(self, device: int) -> int:
        """Create (first call) / look up the plan on ``device``; what the FFI initialize-stage handler does."""
        pl = ctypes.c_void_p()
        _lib.check(_lib.lib().xct_op_plan(self.id, int(device), ctypes.byref(pl)))
        return pl.value

    def apply(self, forward: bool, x, out, device: int, stream: int = 0):
        """``xct_op_apply`` on device pointers (what the FFI execute-stage handler does); ``x`` / ``out`` are
        objects with ``data_ptr()`` and ``numel()`` (e.g. CUDA tensors)."""
        _lib.check(_lib.lib().xct_op_apply(self.id, int(device), 1 if forward else 0, x.data_ptr(), out.data_ptr(),
                                           x.numel(), stream))
        return out

    def release(self):
        self._finalizer()


def _register() -> None:  # pragma: no cover
    global _registered
    if _registered:
        return
    if not HAVE_JAX:
        raise RuntimeError("scico_b200.jax_ffi needs JAX (>= 0.4.35 for jax.ffi)")
    if not os.path.exists(FFI_LIB):
        raise RuntimeError(f"{FFI_LIB} not built: see the header of scico_b200/csrc/xct_ffi.cc")
    lib = ctypes.CDLL(FFI_LIB)
    for name, sym in (("xct_forward", lib.XctForwardFfi), ("xct_adjoint", lib.XctAdjointFfi)):
        # handler bundle: plans are created in the initialize stage (allocation allowed), used in execute
        jax.ffi.register_ffi_target(name, {"initialize": jax.ffi.pycapsule(lib.XctInitFfi), "execute": jax.ffi.pycapsule(sym)},
                                    platform="CUDA")
    _registered = True


def ffi_pair(op):  # pragma: no cover
    """(project, back_project) for the operator ``op`` (:class:`scico_b200.XRayTransform2D` / ``3D``): jittable
    callables, each the linear transpose of the other; leading batch axes (``jax.vmap``) go to the kernels'
    native batch axis (2D) or are looped inside the handler (3D)."""
    _register()
    reg = RegisteredOperator(op)  # captured below: lives as long as the returned callables (and their jit caches)
    nin, nout = len(reg.input_shape), len(reg.output_shape)

    def call(name, core_out, core_in_rank):
        def f(_, a):
            lead = a.shape[: a.ndim - core_in_rank]
            fn = jax.ffi.ffi_call(name, jax.ShapeDtypeStruct(tuple(lead) + tuple(core_out), jnp.float32),
                                  vmap_method="expand_dims")
            return fn(a.astype(jnp.float32), op=np.int64(reg.id))
        return f

    _fwd = call("xct_forward", reg.output_shape, nin)
    _adj = call("xct_adjoint", reg.input_shape, nout)

    # linear_call registers the JVP (the map itself) and the transpose (the other kernel), so jax.grad of
    # 0.5*||A x - y||^2 resolves to the back-projection kernel and jax.linear_transpose(A) works.
    def project(x):
        return linear_call(_fwd, _adj, (), x)

    def back_project(y):
        return linear_call(_adj, _fwd, (), y)

    project.registered_operator = back_project.registered_operator = reg
    return project, back_project
