"""JAX binding of the projector pair (XLA FFI custom calls + linear transpose rules).

Import-guarded: JAX is not installed in the build image, so this module cannot be executed or
tested there; it is kept short and mirrors the in-tree precedent for plugging an external
projector behind ``LinearOperator`` (``scico/linop/xray/astra/_astra_3d.py:498-511``: forward and
adjoint each wrapped so that differentiating one calls the other).

Usage (inside scico, where JAX is present)::

    from scico_b200.jax_ffi import ffi_pair
    proj, bproj = ffi_pair(plan_handle, input_shape, output_shape)
    # XRayTransform3D.__init__:  eval_fn=proj, adj_fn=bproj   (jit / grad / vjp / linear_transpose work)
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

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


def _register() -> None:  # pragma: no cover
    global _registered
    if _registered:
        return
    if not HAVE_JAX:
        raise RuntimeError("scico_b200.jax_ffi needs JAX (>= 0.4.35 for jax.ffi)")
    if not os.path.exists(FFI_LIB):
        raise RuntimeError(f"{FFI_LIB} not built: see the header of scico_b200/csrc/xct_ffi.cc")
    lib = ctypes.CDLL(FFI_LIB)
    jax.ffi.register_ffi_target("xct_forward", jax.ffi.pycapsule(lib.XctForwardFfi), platform="CUDA")
    jax.ffi.register_ffi_target("xct_adjoint", jax.ffi.pycapsule(lib.XctAdjointFfi), platform="CUDA")
    _registered = True


def ffi_pair(plan_handle: int, input_shape, output_shape, batch: int = 1):  # pragma: no cover
    """(project, back_project): jittable callables, each the linear transpose of the other."""
    _register()
    attrs = dict(plan=np.int64(plan_handle), batch=np.int32(batch))
    lead = () if batch == 1 else (batch,)
    fwd_call = jax.ffi.ffi_call("xct_forward", jax.ShapeDtypeStruct(lead + tuple(output_shape), jnp.float32),
                                vmap_method="sequential")
    adj_call = jax.ffi.ffi_call("xct_adjoint", jax.ShapeDtypeStruct(lead + tuple(input_shape), jnp.float32),
                                vmap_method="sequential")

    def _fwd(_, x):
        return fwd_call(x.astype(jnp.float32), **attrs)

    def _adj(_, y):
        return adj_call(y.astype(jnp.float32), **attrs)

    # linear_call registers JVP (the map itself) and transpose (the other kernel), so jax.grad of
    # 0.5*||A x - y||^2 resolves to the back-projection kernel and jax.linear_transpose(A) works.
    def project(x):
        return linear_call(_fwd, _adj, (), x)

    def back_project(y):
        return linear_call(_adj, _fwd, (), y)

    return project, back_project
