"""GPU parity of the fused TV / PDHG kernels (scico_b200/csrc/xct_tv.cuh, through the C ABI) against
the NumPy restatement in oracle/tv_np.py.  Floating-point tolerance: relative L2 <= 1e-5 per kernel
application (same bar as the projectors), <= 1e-4 on the iterate after 25 PDHG iterations."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scico_b200 as sb
from scico_b200 import _lib
from scico_b200.optimize import TVPDHG, FiniteDifference
from oracle import tv_np as T
from oracle import xray_c as C
from oracle import xray_np as O


def _t(torch, a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


@pytest.mark.parametrize("shape", [(5, 6, 7), (1, 1, 1), (2, 33, 65), (16, 16, 16)])
def test_finite_difference_kernels(cuda_device, shape):
    import torch

    rng = np.random.default_rng(0)
    x = rng.standard_normal(shape).astype(np.float32)
    z = rng.standard_normal((3,) + shape).astype(np.float32)
    D = FiniteDifference(shape)
    np.testing.assert_array_equal(D(_t(torch, x, cuda_device)).cpu().numpy(), T.finite_difference(x))
    got = D.adj(_t(torch, z, cuda_device)).cpu().numpy()
    assert O.rel_l2(got, T.finite_difference_adj(z)) <= 1e-6 or np.abs(got - T.finite_difference_adj(z)).max() < 1e-6


def test_finite_difference_slab_halos(cuda_device):
    """Two slabs with halo planes reproduce the unsharded operator exactly."""
    import torch

    rng = np.random.default_rng(1)
    shape = (10, 7, 9)
    x = rng.standard_normal(shape).astype(np.float32)
    z = rng.standard_normal((3,) + shape).astype(np.float32)
    full_f, full_a = T.finite_difference(x), T.finite_difference_adj(z)
    lo, hi = FiniteDifference((6, 7, 9), True, False), FiniteDifference((4, 7, 9), False, True)
    xt, zt = _t(torch, x, cuda_device), _t(torch, z, cuda_device)
    f0 = lo(xt[:6].contiguous(), hi_halo=xt[6].contiguous()).cpu().numpy()
    f1 = hi(xt[6:].contiguous()).cpu().numpy()
    np.testing.assert_array_equal(np.concatenate([f0, f1], axis=1), full_f)
    a0 = lo.adj(zt[:, :6].contiguous()).cpu().numpy()
    a1 = hi.adj(zt[:, 6:].contiguous(), lo_halo=zt[0, 5].contiguous()).cpu().numpy()
    assert np.abs(np.concatenate([a0, a1], axis=0) - full_a).max() < 1e-6


def test_dual_and_primal_kernels_match_oracle(cuda_device):
    import torch

    rng = np.random.default_rng(2)
    shape = (6, 9, 11)
    x = rng.standard_normal(shape).astype(np.float32)
    atz = rng.standard_normal(shape).astype(np.float32)
    z1 = rng.standard_normal((3,) + shape).astype(np.float32)
    z1[:, 2, 3, 4] = 0  # zero-length column: no_nan_divide branch
    tau, sigma, lam, alpha = 0.13, 0.21, 0.4, 1.0
    blk = _lib.TvBlock(*shape, 1, 1)
    L = _lib.lib()
    for nonneg in (0, 1):
        xt, xb = _t(torch, x, cuda_device), torch.empty(shape, device=cuda_device)
        at_t, z1_t = _t(torch, atz, cuda_device), _t(torch, z1, cuda_device)  # keep alive across the launch
        _lib.check(L.xct_tv_primal_step(ctypes.byref(blk), xt.data_ptr(), xb.data_ptr(), at_t.data_ptr(),
                                        z1_t.data_ptr(), None, tau, alpha, nonneg, None))
        torch.cuda.synchronize()
        arg = x - np.float32(tau) * (atz + T.finite_difference_adj(z1))
        want = np.maximum(arg, 0) if nonneg else arg
        assert O.rel_l2(xt.cpu().numpy(), want) <= 1e-5
        assert O.rel_l2(xb.cpu().numpy(), 2 * want - x) <= 1e-5
    xbar = rng.standard_normal(shape).astype(np.float32)
    xbar[2, 3, 4:6] = xbar[2, 3, 4]
    zt, xb_t = _t(torch, z1, cuda_device), _t(torch, xbar, cuda_device)
    _lib.check(L.xct_tv_dual_step(ctypes.byref(blk), zt.data_ptr(), xb_t.data_ptr(), None, sigma, lam, None))
    torch.cuda.synchronize()
    p = z1 + np.float32(sigma) * T.finite_difference(xbar)
    want = T.conj_prox(lambda v, l: T.l21_prox(v, np.float32(lam) * np.float32(l)), p, sigma)
    assert O.rel_l2(zt.cpu().numpy(), want) <= 1e-5
    assert np.all(np.isfinite(zt.cpu().numpy()))
    # every column of the dual variable lies in the lam-ball
    assert float(torch.sqrt((zt ** 2).sum(dim=0)).max()) <= lam * (1 + 1e-5)
    y, ax, z0 = (rng.standard_normal(50).astype(np.float32) for _ in range(3))
    z0t, ax_t, y_t = _t(torch, z0, cuda_device), _t(torch, ax, cuda_device), _t(torch, y, cuda_device)
    _lib.check(L.xct_l2_dual_step(50, z0t.data_ptr(), ax_t.data_ptr(), y_t.data_ptr(), sigma, None))
    torch.cuda.synchronize()
    want = T.conj_prox(lambda v, l: T.sql2_prox(v, y, l), z0 + np.float32(sigma) * ax, sigma)
    assert O.rel_l2(z0t.cpu().numpy(), want) <= 1e-5


@pytest.mark.parametrize("nonneg", [False, True])
def test_pdhg_iterations_match_oracle(cuda_device, nonneg):
    """25 iterations of TVPDHG (CUDA projector pair + fused kernels) vs the oracle's PDHG with the
    oracle's projectors on the same problem."""
    import torch

    N, D, V = (8, 24, 20), (8, 32), 10
    M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
    A = sb.XRayTransform3D(N, M, D)
    x_gt = np.zeros(N, np.float32)
    x_gt[2:6, 6:16, 5:14] = 1.0
    x_gt[3:5, 9:12, 8:11] = 2.0
    Ao = lambda x: C.project_3d(x, A.matrices, D)
    ATo = lambda y: C.back_project_3d(y, A.matrices, N)
    y = Ao(x_gt) + 0.05 * np.random.default_rng(3).standard_normal((V,) + D).astype(np.float32)
    lam, tau, sigma = 0.1, 0.05, 0.05
    S = TVPDHG(A, _t(torch, y, cuda_device), lam, tau, sigma, nonneg=nonneg, maxiter=25)
    x, z0, z1 = np.zeros(N, np.float32), np.zeros_like(y), np.zeros((3,) + N, np.float32)
    for _ in range(25):
        x, z0, z1 = T.pdhg_tv_step(x, z0, z1, Ao, ATo, y, lam, tau, sigma, nonneg=nonneg)
    S.solve()
    assert O.rel_l2(S.x.cpu().numpy(), x) <= 1e-4
    assert O.rel_l2(S.z1.cpu().numpy(), z1) <= 1e-4
    assert O.rel_l2(S.z0.cpu().numpy(), z0) <= 1e-4
    assert abs(S.objective() - T.tv_objective(x, Ao, y, lam)) <= 1e-4 * T.tv_objective(x, Ao, y, lam)


def test_pdhg_itstats_and_parameter_estimate(cuda_device):
    import torch

    N, D, V = (8, 24, 20), (8, 32), 10
    M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
    A = sb.XRayTransform3D(N, M, D)
    tau, sigma = TVPDHG.estimate_parameters(A, factor=0.9, maxiter=30)
    # ||C||^2 by the same power iteration on the oracle
    v = np.random.default_rng(0).standard_normal(N).astype(np.float32)
    for _ in range(60):
        v /= np.linalg.norm(v)
        w = C.back_project_3d(C.project_3d(v, A.matrices, D), A.matrices, N) + T.finite_difference_adj(T.finite_difference(v))
        mu = float(np.sum(v * w))
        v = w
    assert abs(tau * sigma * mu - 0.9) < 0.05
    x_gt = np.zeros(N, np.float32)
    x_gt[2:6, 6:16, 5:14] = 1.0
    y = _t(torch, C.project_3d(x_gt, A.matrices, D), cuda_device)
    S = TVPDHG(A, y, 0.05, tau, sigma, maxiter=60, itstat=True)
    S.solve()
    obj = [h["objective"] for h in S.history]
    assert obj[-1] < 0.02 * S.objective(torch.zeros(N, device=cuda_device))
    # the first iteration starts from zero duals and leaves x unchanged: compare with the second
    assert S.history[0]["prml_rsdl"] == 0.0 and S.history[-1]["prml_rsdl"] < S.history[1]["prml_rsdl"]
    assert O.rel_l2(S.x.cpu().numpy(), x_gt) < 0.15


@pytest.mark.parametrize("alpha,nonneg,with_x0", [(1.0, False, False), (0.5, True, True), (0.0, False, True)])
def test_pdhg_fused_iteration_statistics_match_explicit_ones(cuda_device, alpha, nonneg, with_x0):
    """The statistics the iteration's own kernels accumulate (xct_*_step_stat, xct_tv_norm; A x by the recurrence
    A x_new = (A xbar + alpha A x_old) / (1 + alpha)) equal the reference's definitions evaluated explicitly
    (_primaldual.py:175-217): objective with its own forward projection, residuals from copies of the old iterates;
    and the iterates are those of the solver without statistics, bit for bit."""
    import torch

    N, D, V = (8, 24, 20), (8, 32), 10
    M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
    A = sb.XRayTransform3D(N, M, D)
    rng = np.random.default_rng(5)
    x_gt = np.zeros(N, np.float32)
    x_gt[2:6, 6:16, 5:14] = 1.0
    y = _t(torch, C.project_3d(x_gt, A.matrices, D) + 0.05 * rng.standard_normal((V,) + D).astype(np.float32), cuda_device)
    x0 = _t(torch, rng.random(N).astype(np.float32), cuda_device) if with_x0 else None
    lam, tau, sigma = 0.1, 0.05, 0.05
    S = TVPDHG(A, y, lam, tau, sigma, alpha=alpha, nonneg=nonneg, x0=x0, maxiter=30, itstat=True)
    P = TVPDHG(A, y, lam, tau, sigma, alpha=alpha, nonneg=nonneg, x0=x0, maxiter=30)
    for it in range(30):
        xo, z0o, z1o = S.x.clone(), S.z0.clone(), S.z1.clone()
        S.step()
        P.step()
        h = S.history[-1]
        assert h["iter"] == it + 1
        obj = S.objective()
        assert abs(h["objective"] - obj) <= 2e-6 * abs(obj)
        pr = float((S.x - xo).double().norm()) / tau
        du = float(torch.sqrt((S.z0 - z0o).double().norm() ** 2 + (S.z1 - z1o).double().norm() ** 2)) / sigma
        assert abs(h["prml_rsdl"] - pr) <= 1e-9 * max(pr, 1e-30) + 1e-12
        assert abs(h["dual_rsdl"] - du) <= 1e-9 * max(du, 1e-30) + 1e-12
    assert torch.equal(S.x, P.x) and torch.equal(S.z0, P.z0) and torch.equal(S.z1, P.z1)
    assert O.rel_l2(S.ax_x.cpu().numpy(), A.project(S.x).cpu().numpy()) <= 1e-6


@pytest.mark.gpu
def test_solve_replays_a_cuda_graph_and_matches_plain_stepping(cuda_device):
    """solve() captures one iteration in a CUDA graph when nothing in it needs the host (PDHG and the
    split ADMM variants, single GPU, statistics off) and replays it: same iterates as stepping."""
    import torch

    import scico_b200 as sb
    from scico_b200.optimize import TVADMM, TVPDHG, TVProximalADMM

    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(3)
    A2 = sb.XRayTransform2D((64, 56), np.linspace(0, np.pi, 20, endpoint=False))
    N, D = (16, 40, 48), (16, 64)
    A3 = sb.XRayTransform3D(N, sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, 9, endpoint=False)[:, None]), D)
    for A in (A2, A3):
        x_gt = torch.rand(A.input_shape, device=dev, generator=g)
        y = A(x_gt)
        for make in (lambda: TVPDHG(A, y, 0.1, 0.02, 0.02, maxiter=12),
                     lambda: TVProximalADMM(A, y, lam=0.5, rho=0.1, mu=2e3, nu=1.01, alpha=2.0, maxiter=12)):
            S, R = make(), make()
            S.solve(use_graph=True)
            for _ in range(12):
                R.step()
            assert S.itnum == R.itnum == 12
            rel = (torch.linalg.vector_norm(S.x - R.x) / torch.linalg.vector_norm(R.x)).item()
            assert rel <= 1e-5, rel
            E = make()
            E.solve()  # default: plain stepping
            assert (torch.linalg.vector_norm(E.x - R.x) / torch.linalg.vector_norm(R.x)).item() <= 1e-5
    with pytest.raises(ValueError):
        TVADMM(A2, A2(torch.rand(A2.input_shape, device=dev)), 0.5, 5.0, maxiter=3).solve(use_graph=True)
