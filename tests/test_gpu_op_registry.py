"""The operator registry behind the XLA FFI handlers (xct_op_*, include/scico_b200_xray.h) on the device: what
scico_b200/csrc/xct_ffi.cc does per custom call, driven through ctypes because JAX is not installable here --
plan creation per device in the "initialize" step, application by operator id with the batch folded from the
element count (jax.vmap's expand_dims rule), lifetime / released ids."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scico_b200 as sb
from scico_b200 import _lib
from scico_b200.jax_ffi import RegisteredOperator
from oracle import xray_c as C
from oracle import xray_np as O


def test_registered_operator_applies_by_id_with_a_batch(cuda_device):
    import torch

    rng = np.random.default_rng(3)
    # 2D: native batch axis
    nx, V = (40, 36), 24
    A = sb.XRayTransform2D(nx, np.linspace(0, np.pi, V, endpoint=False))
    reg = RegisteredOperator(A)
    with pytest.raises(_lib.XctError):  # execute before initialize: no plan on this device yet, and no allocation
        reg.apply(True, torch.zeros((1,) + nx, device=cuda_device), torch.zeros((1, V, A.ny), device=cuda_device), 0)
    assert reg.plan(0) and reg.plan(0) == reg.plan(0)  # created once per device
    x = rng.standard_normal((2, 3) + nx).astype(np.float32)  # two leading batch axes (nested vmap)
    xt = torch.as_tensor(x, device=cuda_device)
    out = torch.empty((2, 3, V, A.ny), device=cuda_device)
    reg.apply(True, xt, out, 0, torch.cuda.current_stream().cuda_stream)
    T = A.view_table
    for i in range(2):
        for j in range(3):
            assert O.rel_l2(out[i, j].cpu().numpy(), C.project_2d(x[i, j], T, A.ny)) <= 1e-5
    back = torch.empty((2, 3) + nx, device=cuda_device)
    reg.apply(False, out, back, 0, torch.cuda.current_stream().cuda_stream)
    assert O.rel_l2(back[1, 2].cpu().numpy(), C.back_project_2d(out[1, 2].cpu().numpy(), T, nx)) <= 1e-5
    with pytest.raises(_lib.XctError):  # operand that is not a whole number of images
        reg.apply(True, xt.reshape(-1)[:-1], out, 0)
    # 3D: batch items run one after the other inside the call
    N, D, W = (12, 20, 24), (12, 32), 6
    M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, W, endpoint=False)[:, None])
    B = sb.XRayTransform3D(N, M, D)
    r3 = RegisteredOperator(B)
    r3.plan(0)
    v = rng.standard_normal((3,) + N).astype(np.float32)
    o3 = torch.empty((3, W) + D, device=cuda_device)
    r3.apply(True, torch.as_tensor(v, device=cuda_device), o3, 0, torch.cuda.current_stream().cuda_stream)
    for i in range(3):
        assert O.rel_l2(o3[i].cpu().numpy(), C.project_3d(v[i], B.matrices, D)) <= 1e-5
    # lifetime: the id outlives the Python operator it was made from, and dies with the last reference
    oid = r3.id
    del B
    r3.apply(True, torch.as_tensor(v[:1], device=cuda_device), o3[:1], 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    r3.release()
    assert _lib.lib().xct_op_release(oid) == _lib.XCT_ERR_INVALID  # already gone: an error, not a crash
    with pytest.raises(_lib.XctError):
        r3.apply(True, torch.as_tensor(v[:1], device=cuda_device), o3[:1], 0)
