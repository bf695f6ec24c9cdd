"""torch.autograd through the native pair on the B200: the checks the reference applies to an external
projector behind LinearOperator (scico/test/linop/xray/astra/test_astra_2d.py:117-187) -- <x, A^T A x> =
||A x||^2, <y, A A^T y> = ||A^T y||^2, grad ||A x||^2 = 2 A^T A x, gradient through A.T.  The backward pass
must be the other kernel of the pair (scico_b200/xray.py::_ProjectorFn)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scico_b200 as sb
from scico_b200 import _lib


def _ops():
    N, D, V = (12, 40, 36), (12, 56), 9
    M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
    return [sb.XRayTransform2D((48, 40), np.linspace(0, np.pi, 20, endpoint=False)), sb.XRayTransform3D(N, M, D)]


def _rel(torch, a, b):
    return (torch.linalg.vector_norm(a.double() - b.double()) / torch.linalg.vector_norm(b.double())).item()


@pytest.mark.parametrize("which", [0, 1])
def test_gram_identities(cuda_device, which):
    import torch

    A = _ops()[which]
    g = torch.Generator(device=cuda_device).manual_seed(3)
    x = torch.randn(A.input_shape, device=cuda_device, generator=g)
    y = torch.randn(A.output_shape, device=cuda_device, generator=g)
    Ax, ATy = A(x), A.T(y)
    lhs = (x.double() * A.T(Ax).double()).sum().item()
    assert abs(lhs - (Ax.double() ** 2).sum().item()) <= 1e-5 * lhs
    lhs = (y.double() * A(ATy).double()).sum().item()
    assert abs(lhs - (ATy.double() ** 2).sum().item()) <= 1e-5 * lhs


@pytest.mark.parametrize("which", [0, 1])
def test_grad_is_the_adjoint_kernel(cuda_device, which):
    import torch

    A = _ops()[which]
    g = torch.Generator(device=cuda_device).manual_seed(4)
    x = torch.randn(A.input_shape, device=cuda_device, generator=g).requires_grad_()
    y = torch.randn(A.output_shape, device=cuda_device, generator=g).requires_grad_()

    _lib.launch_count_reset()
    loss = (A(x) ** 2).sum()
    assert _lib.launch_count() > 0  # (the counter is per thread: autograd's worker runs the backward launches)
    (gx,) = torch.autograd.grad(loss, x)
    with torch.no_grad():
        want = 2 * A.adj(A(x))
    assert gx.shape == x.shape and _rel(torch, gx, want) <= 1e-5

    (gy,) = torch.autograd.grad((A.T(y) ** 2).sum(), y)  # gradient through the transpose
    with torch.no_grad():
        want = 2 * A(A.adj(y))
    assert gy.shape == y.shape and _rel(torch, gy, want) <= 1e-5

    # second derivative of 1/2 ||A x||^2 along c is A^T A c
    c = torch.randn(A.input_shape, device=cuda_device, generator=g)
    (g1,) = torch.autograd.grad(0.5 * (A(x) ** 2).sum(), x, create_graph=True)
    (g2,) = torch.autograd.grad((g1 * c).sum(), x)
    with torch.no_grad():
        assert _rel(torch, g2, A.adj(A(c))) <= 1e-5
        assert not A(x).requires_grad
