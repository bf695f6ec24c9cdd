"""GPU parity of the ADMM-family kernels and drivers (scico_b200/csrc/xct_solver.cuh through the C ABI,
scico_b200/optimize.py) against the NumPy restatement in oracle/tv_np.py.  Tolerances (floating
point; written here as the task requires): relative L2 <= 1e-5 per kernel application, <= 1e-4 on the
iterates after a fixed number of solver iterations (CG runs a fixed number of inner iterations there so
that both sides do the same work; a second test uses the reference's tolerance-based stop)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scico_b200 as sb
from scico_b200 import _lib
from scico_b200.optimize import TVADMM, TVLinearizedADMM, TVProximalADMM
from oracle import tv_np as T
from oracle import xray_c as C
from oracle import xray_np as O

f32 = np.float32


def _t(torch, a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


def _blk(shape, first=1, last=1):
    return ctypes.byref(_lib.TvBlock(*shape, first, last))


def _close(got, want, tol=1e-5):
    got = got.detach().cpu().numpy() if hasattr(got, "detach") else got
    return O.rel_l2(got, want) <= tol or np.abs(got - want).max() <= 1e-6


@pytest.mark.parametrize("shape", [(6, 9, 11), (1, 17, 13), (3, 1, 5)])
def test_grad_prox_and_sino_prox_kernels(cuda_device, shape):
    import torch

    L = _lib.lib()
    rng = np.random.default_rng(0)
    x = rng.standard_normal(shape).astype(f32)
    z = rng.standard_normal((3,) + shape).astype(f32)
    u = rng.standard_normal((3,) + shape).astype(f32)
    u[:, 0, 0, 0] = 0
    x[0, 0, :2] = x[0, 0, 0]  # a zero-length column for the ADMM / LADMM argument (no_nan_divide branch)
    dscale, thr, inv_nu = 1.7, 0.6, 1 / 1.01
    Dx = T.finite_difference(x)
    xt = _t(torch, x, cuda_device)
    # ADMM / LADMM (dscale = 1)
    for mode in (_lib.SPLIT_ADMM, _lib.SPLIT_LADMM):
        zt, ut, wt = _t(torch, z, cuda_device), _t(torch, u, cuda_device), torch.zeros((3,) + shape, device=cuda_device)
        _lib.check(L.xct_grad_prox_step(_blk(shape), xt.data_ptr(), None, zt.data_ptr(), ut.data_ptr(),
                                        wt.data_ptr() if mode else None, 1.0, thr, 1.0, mode, None))
        torch.cuda.synchronize()
        zn = T.l21_prox(Dx + u, thr)
        un = u + Dx - zn
        assert _close(zt, zn) and _close(ut, un)
        assert np.all(np.isfinite(zt.cpu().numpy()))
        if mode == _lib.SPLIT_LADMM:
            assert _close(wt, (Dx - zn) + un)
    # PADMM
    zt, ut, wt = _t(torch, z, cuda_device), _t(torch, u, cuda_device), torch.zeros((3,) + shape, device=cuda_device)
    _lib.check(L.xct_grad_prox_step(_blk(shape), xt.data_ptr(), None, zt.data_ptr(), ut.data_ptr(), wt.data_ptr(),
                                    dscale, thr, inv_nu, _lib.SPLIT_PADMM, None))
    torch.cuda.synchronize()
    Cx = f32(dscale) * Dx
    zn = T.l21_prox(z + f32(inv_nu) * ((Cx - z) + u), thr)
    un = (u + Cx) - zn
    assert _close(zt, zn) and _close(ut, un) and _close(wt, 2 * un - u)

    n = 77
    ax, y, z0, u0 = (rng.standard_normal(n).astype(f32) for _ in range(4))
    for mode, c in ((_lib.SPLIT_LADMM, 0.8), (_lib.SPLIT_PADMM, 1 / (5e-3 * 1.01))):
        a_t, y_t, z_t, u_t = (_t(torch, a, cuda_device) for a in (ax, y, z0, u0))
        w_t = torch.zeros(n, device=cuda_device)
        _lib.check(L.xct_sino_prox_step(n, a_t.data_ptr(), y_t.data_ptr(), z_t.data_ptr(), u_t.data_ptr(),
                                        w_t.data_ptr(), c, inv_nu, mode, None))
        torch.cuda.synchronize()
        v = ax + u0 if mode == _lib.SPLIT_LADMM else z0 + f32(inv_nu) * ((ax - z0) + u0)
        zn = T.sql2_prox(v, y, c)
        un = (u0 + ax) - zn
        wn = (ax - zn) + un if mode == _lib.SPLIT_LADMM else 2 * un - u0
        assert _close(z_t, zn) and _close(u_t, un) and _close(w_t, wn)


@pytest.mark.parametrize("shape", [(6, 9, 11), (1, 17, 13)])
def test_primal_rhs_and_cg_kernels(cuda_device, shape):
    import torch

    L = _lib.lib()
    rng = np.random.default_rng(1)
    x, atq, aty, b = (rng.standard_normal(shape).astype(f32) for _ in range(4))
    w, z, u = (rng.standard_normal((3,) + shape).astype(f32) for _ in range(3))
    n = x.size
    step, dscale, rho = 0.07, 3.0, 5.0
    for nonneg in (0, 1):
        xt, a_t, w_t = _t(torch, x, cuda_device), _t(torch, atq, cuda_device), _t(torch, w, cuda_device)
        _lib.check(L.xct_grad_primal_step(_blk(shape), xt.data_ptr(), a_t.data_ptr(), w_t.data_ptr(), None,
                                          step, dscale, nonneg, None))
        torch.cuda.synchronize()
        arg = x - f32(step) * (atq + f32(dscale) * T.finite_difference_adj(w))
        assert _close(xt, np.maximum(arg, 0) if nonneg else arg)
    # right-hand side and its squared norm
    sc = torch.zeros(8, dtype=torch.float64, device=cuda_device)
    sp = lambda i: sc.data_ptr() + 8 * i  # noqa: E731
    aty_t, z_t, u_t = _t(torch, aty, cuda_device), _t(torch, z, cuda_device), _t(torch, u, cuda_device)
    rhs_t = torch.empty(shape, device=cuda_device)
    _lib.check(L.xct_admm_rhs(_blk(shape), aty_t.data_ptr(), z_t.data_ptr(), u_t.data_ptr(), None, rho,
                              rhs_t.data_ptr(), sp(5), None))
    torch.cuda.synchronize()
    rhs = aty + f32(rho) * T.finite_difference_adj(z - u)
    assert _close(rhs_t, rhs)
    assert abs(float(sc[5]) - float(np.sum(rhs.astype(np.float64) ** 2))) <= 1e-6 * float(np.sum(rhs.astype(np.float64) ** 2))
    # CG start, one full CG iteration
    dtd = lambda v: T.finite_difference_adj(T.finite_difference(v))  # noqa: E731
    atax, atap = (rng.standard_normal(shape).astype(f32) for _ in range(2))
    xt, b_t, atax_t, atap_t = (_t(torch, a, cuda_device) for a in (x, b, atax, atap))
    r_t, p_t, q_t = (torch.empty(shape, device=cuda_device) for _ in range(3))
    _lib.check(L.xct_cg_init(_blk(shape), xt.data_ptr(), None, None, atax_t.data_ptr(), b_t.data_ptr(), rho,
                             r_t.data_ptr(), p_t.data_ptr(), sp(0), None))
    torch.cuda.synchronize()
    r = b - (f32(rho) * dtd(x) + atax)
    assert _close(r_t, r) and _close(p_t, r)
    num = float(np.sum(r.astype(np.float64) ** 2))
    assert abs(float(sc[0]) - num) <= 1e-6 * num
    sc[1] = 123.0  # the ring slot the lhs kernel must clear
    _lib.check(L.xct_cg_lhs(_blk(shape), p_t.data_ptr(), None, None, atap_t.data_ptr(), rho, q_t.data_ptr(), sp(3), sp(1), None))
    torch.cuda.synchronize()
    q = f32(rho) * dtd(r) + atap
    assert _close(q_t, q)
    pq = float(np.sum(r.astype(np.float64) * q))
    assert abs(float(sc[3]) - pq) <= 1e-5 * abs(pq) + 1e-9 and float(sc[1]) == 0.0
    sc[4] = 7.0
    _lib.check(L.xct_cg_update_xr(n, xt.data_ptr(), r_t.data_ptr(), p_t.data_ptr(), q_t.data_ptr(), sp(0), sp(3), sp(1), sp(4), None))
    torch.cuda.synchronize()
    alpha = f32(num / pq)
    xn, rn = x + alpha * r, r - alpha * q
    assert _close(xt, xn) and _close(r_t, rn, 1e-4)
    num_new = float(np.sum(rn.astype(np.float64) ** 2))
    assert abs(float(sc[1]) - num_new) <= 1e-3 * num_new + 1e-9 and float(sc[4]) == 0.0
    _lib.check(L.xct_cg_update_p(n, p_t.data_ptr(), r_t.data_ptr(), sp(0), sp(1), None))
    torch.cuda.synchronize()
    assert _close(p_t, rn + f32(num_new / num) * r, 1e-4)


def test_solver_kernels_slab_halos(cuda_device):
    """Two z-slabs with halo planes reproduce the unsharded kernels (what SlabSharded solvers rely on)."""
    import torch

    L = _lib.lib()
    rng = np.random.default_rng(2)
    shape, cut = (10, 7, 9), 6
    top, bot = (cut,) + shape[1:], (shape[0] - cut,) + shape[1:]
    p, atap, aty = (rng.standard_normal(shape).astype(f32) for _ in range(3))
    z, u = (rng.standard_normal((3,) + shape).astype(f32) for _ in range(2))
    rho = 2.5
    dtd = T.finite_difference_adj(T.finite_difference(p))
    q_full = f32(rho) * dtd + atap
    rhs_full = aty + f32(rho) * T.finite_difference_adj(z - u)
    pt, at, yt, zt, ut = (_t(torch, a, cuda_device) for a in (p, atap, aty, z, u))
    sc = torch.zeros(8, dtype=torch.float64, device=cuda_device)
    sp = lambda i: sc.data_ptr() + 8 * i  # noqa: E731
    q0, q1 = torch.empty(top, device=cuda_device), torch.empty(bot, device=cuda_device)
    p0, p1 = pt[:cut].contiguous(), pt[cut:].contiguous()
    a0, a1 = at[:cut].contiguous(), at[cut:].contiguous()
    hi0, lo1 = pt[cut].contiguous(), pt[cut - 1].contiguous()
    _lib.check(L.xct_cg_lhs(_blk(top, 1, 0), p0.data_ptr(), None, hi0.data_ptr(), a0.data_ptr(), rho, q0.data_ptr(), sp(3), None, None))
    _lib.check(L.xct_cg_lhs(_blk(bot, 0, 1), p1.data_ptr(), lo1.data_ptr(), None, a1.data_ptr(), rho, q1.data_ptr(), sp(3), None, None))
    torch.cuda.synchronize()
    got = torch.cat([q0, q1]).cpu().numpy()
    assert np.abs(got - q_full).max() <= 1e-5
    pq = float(np.sum(p.astype(np.float64) * q_full))
    assert abs(float(sc[3]) - pq) <= 1e-5 * abs(pq) + 1e-6  # both slabs accumulate into one scalar
    r0, r1 = torch.empty(top, device=cuda_device), torch.empty(bot, device=cuda_device)
    z0, z1 = zt[:, :cut].contiguous(), zt[:, cut:].contiguous()
    u0, u1 = ut[:, :cut].contiguous(), ut[:, cut:].contiguous()
    y0, y1 = yt[:cut].contiguous(), yt[cut:].contiguous()
    lo_zu = (zt[0, cut - 1] - ut[0, cut - 1]).contiguous()
    _lib.check(L.xct_admm_rhs(_blk(top, 1, 0), y0.data_ptr(), z0.data_ptr(), u0.data_ptr(), None, rho, r0.data_ptr(), sp(5), None))
    _lib.check(L.xct_admm_rhs(_blk(bot, 0, 1), y1.data_ptr(), z1.data_ptr(), u1.data_ptr(), lo_zu.data_ptr(), rho, r1.data_ptr(), sp(5), None))
    torch.cuda.synchronize()
    assert np.abs(torch.cat([r0, r1]).cpu().numpy() - rhs_full).max() <= 1e-5
    # grad_prox with the next slab's first plane as halo
    x = rng.standard_normal(shape).astype(f32)
    xt = _t(torch, x, cuda_device)
    zf, uf = zt.clone(), ut.clone()
    _lib.check(L.xct_grad_prox_step(_blk(shape), xt.data_ptr(), None, zf.data_ptr(), uf.data_ptr(), None, 1.0, 0.5, 1.0, 0, None))
    x0, hi = xt[:cut].contiguous(), xt[cut].contiguous()
    _lib.check(L.xct_grad_prox_step(_blk(top, 1, 0), x0.data_ptr(), hi.data_ptr(), z0.data_ptr(), u0.data_ptr(), None, 1.0, 0.5, 1.0, 0, None))
    torch.cuda.synchronize()
    assert torch.equal(z0, zf[:, :cut]) and torch.equal(u0, uf[:, :cut])


def _problem_3d():
    N, D, V = (8, 24, 20), (8, 32), 10
    M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
    A = sb.XRayTransform3D(N, M, D)
    x_gt = np.zeros(N, f32)
    x_gt[2:6, 6:16, 5:14] = 1.0
    x_gt[3:5, 9:12, 8:11] = 2.0
    Ao = lambda x: C.project_3d(x, A.matrices, D)
    ATo = lambda y: C.back_project_3d(y, A.matrices, N)
    y = Ao(x_gt) + 0.05 * np.random.default_rng(3).standard_normal((V,) + D).astype(f32)
    return N, A, Ao, ATo, x_gt, y


def _problem_2d():
    nx, V = (40, 36), 24
    angles = np.linspace(0, np.pi, V, endpoint=False)
    A = sb.XRayTransform2D(nx, angles)
    Tb = O.view_table_2d(angles, A.x0, A.dx, A.y0)
    Ao = lambda x: C.project_2d(x, Tb, A.ny)
    ATo = lambda y: C.back_project_2d(y, Tb, nx)
    x_gt = np.zeros(nx, f32)
    x_gt[10:30, 8:25] = 1.0
    x_gt[15:22, 12:18] = 0.3
    y = Ao(x_gt) + 0.05 * np.random.default_rng(4).standard_normal((V, A.ny)).astype(f32)
    return nx, A, Ao, ATo, x_gt, y


@pytest.mark.parametrize("dim", [2, 3])
def test_admm_iterations_match_oracle(cuda_device, dim):
    """ADMM + CG (ct_tv_admm.py set-up: lam = 2, rho = 5) with a fixed number of CG iterations per
    x-step: x, z, u after 6 iterations vs the oracle's ADMM on the oracle's projectors."""
    import torch

    N, A, Ao, ATo, x_gt, y = _problem_2d() if dim == 2 else _problem_3d()
    lam, rho, iters, cg_it = 2.0, 5.0, 6, 5
    x0 = np.clip(ATo(y) / f32(ATo(Ao(np.ones(N, f32))).max()), 0, 1).astype(f32)
    S = TVADMM(A, _t(torch, y, cuda_device), lam, rho, x0=_t(torch, x0, cuda_device), maxiter=iters,
               cg_tol=1e-30, cg_maxiter=cg_it)
    x, z, u = T.admm_tv_init(x0)
    for _ in range(iters):
        x, z, u, info = T.admm_tv_step(x, z, u, Ao, ATo, y, lam, rho, cg_tol=1e-30, cg_maxiter=cg_it)
        assert info["num_iter"] == cg_it
    S.solve()
    assert S.cg_info["num_iter"] == cg_it and S.cg_iters_total == iters * cg_it
    assert O.rel_l2(S.x.cpu().numpy(), x) <= 1e-4
    assert O.rel_l2(S.z.cpu().numpy(), z) <= 1e-4
    assert O.rel_l2(S.u.cpu().numpy(), u) <= 1e-4
    assert tuple(S.z.shape) == (dim,) + tuple(N)
    if dim == 2:  # the axis-0 component of the (1, N0, N1) block never leaves zero
        assert float(S.z1[0].abs().max()) == 0.0 and float(S.u1[0].abs().max()) == 0.0


def test_admm_with_tolerance_stop_converges_like_the_oracle(cuda_device):
    """Reference defaults (cg tol 1e-4, maxiter 25, ct_tv_admm.py:64-66): the stop test can fire one CG
    iteration apart on the two sides, so the iterates agree to the CG tolerance, not to 1e-5."""
    import torch

    N, A, Ao, ATo, x_gt, y = _problem_2d()
    lam, rho, iters = 0.5, 5.0, 15
    S = TVADMM(A, _t(torch, y, cuda_device), lam, rho, maxiter=iters, cg_tol=1e-4, cg_maxiter=25, itstat=True)
    x, z, u = T.admm_tv_init(np.zeros(N, f32))
    its = []
    for _ in range(iters):
        x, z, u, info = T.admm_tv_step(x, z, u, Ao, ATo, y, lam, rho, cg_tol=1e-4, cg_maxiter=25)
        its.append(info["num_iter"])
    def check_objective(solver):
        # the statistics' objective uses A x carried through the CG updates (A x += alpha A p): equal to the objective
        # evaluated with a forward projection of its own
        zt = solver.z1.double()
        explicit = 0.5 * float(((A.project(solver.x) - solver.y).double() ** 2).sum()) + lam * float(torch.sqrt((zt ** 2).sum(0)).sum())
        assert abs(solver.history[-1]["objective"] - explicit) <= 1e-5 * explicit

    S.solve(callback=check_objective)
    assert abs(S.cg_iters_total - sum(its)) <= iters
    assert O.rel_l2(S.x.cpu().numpy(), x) <= 5e-3  # one CG iteration apart moves x by a few 1e-3 (seen: 2.05e-3)
    h = S.history
    assert len(h) == iters and h[-1]["prml_rsdl"] < h[0]["prml_rsdl"] and all(r["cg_rel_res"] <= 1.01e-4 or r["cg_iters"] == 25 for r in h)
    obj = T.tv_objective(x, Ao, y, lam)
    assert abs(S.objective() - obj) <= 5e-3 * obj  # same reason (seen: 2.4e-3)
    assert O.rel_l2(S.x.cpu().numpy(), x_gt) < 0.25


def _disc_phantom(n, seed=1234, discs=40):
    """Seeded union of discs (stand-in for xdesign's Foam of examples/scripts/ct_tv_admm.py:40-46)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n].astype(f32) / n - 0.5
    img = ((xx ** 2 + yy ** 2) < 0.45 ** 2).astype(f32)
    for _ in range(discs):
        cx, cy = rng.uniform(-0.3, 0.3, 2)
        r = rng.uniform(0.01, 0.06)
        img[(xx - cx) ** 2 + (yy - cy) ** 2 < r * r] = 0.0
    return img


def _cg_fp64_sums(Aop, b, x0, k):
    """k iterations of scico.solver.cg's recurrences (oracle/tv_np.py::cg) with the inner products accumulated
    in fp64 -- the device's accumulation -- and everything else in fp32: a second, equally valid arithmetic."""
    x = np.asarray(x0, f32)
    r = (b - Aop(x)).astype(f32)
    p = r
    num = np.sum(r.astype(np.float64) ** 2)
    for _ in range(k):
        Ap = Aop(p)
        alpha = f32(num / np.sum(p.astype(np.float64) * Ap))
        x = (x + alpha * p).astype(f32)
        r = (r - alpha * Ap).astype(f32)
        num_old, num = num, np.sum(r.astype(np.float64) ** 2)
        p = (r + f32(num / num_old) * p).astype(f32)
    return x


def test_c1_tv_admm_reference_defaults_every_iteration_matches_oracle(cuda_device):
    """BASELINE.json configs[0] (C1: 256^2, 180 views) with the reference example's parameters
    (ct_tv_admm.py:61-84: lam 2, rho 5, 25 iterations, CG tol 1e-4 / maxiter 25, x0 = clip(fbp)).

    Free-running, two correct fp32 implementations of this solve do NOT agree to 1e-4: measured with the
    ORACLE ALONE, accumulating CG's inner products in fp32 versus fp64 changes the result by 2.3e-2 and the
    total CG count from 106 to 128 (same order as the GPU-vs-oracle gap of round 1, 2.5e-2 / 130 vs 188).
    Two mechanisms: (i) the x-step is warm-started, so <r, r> at entry sits within a few per cent of the
    threshold (0.82 - 1.04 of it in half of the iterations) and CG runs either 0 or 15 - 22 iterations on a
    coin flip of the last bits; (ii) fp32 CG amplifies a 1e-7 perturbation by ~2.5x per iteration (one x-step
    from identical state: 6e-9 after 1, 6e-8 after 5, 5e-6 after 15, 4e-4 after 25 iterations).

    What CAN be pinned, and is: every one of the 25 ADMM iterations, started from the GPU solver's own state,
    equals the oracle's iteration from that state with the same CG depth (tolerance: 1e-4 up to depth 10, then
    doubling per CG iteration -- the measured amplification -- or 8x the oracle's own sensitivity to
    rounding-level changes of its arithmetic at that state and depth, whichever is larger, capped at 2e-2); the <r, r> sequence the stop test reads agrees with the oracle's; the stop test is
    the reference's fp32 expression; and the reconstruction quality equals a free-running oracle's."""
    import torch

    n, V = 256, 180
    angles = np.linspace(0, np.pi, V, endpoint=False)
    A = sb.XRayTransform2D((n, n), angles)
    Tb = O.view_table_2d(angles, A.x0, A.dx, A.y0)
    Ao = lambda v: C.project_2d(v, Tb, A.ny)  # noqa: E731
    ATo = lambda v: C.back_project_2d(v, Tb, (n, n))  # noqa: E731
    x_gt = _disc_phantom(n)
    y = A(_t(torch, x_gt, cuda_device))
    x0 = torch.clamp(A.fbp(y), 0.0, 1.0)
    yn, x0n = y.cpu().numpy(), x0.cpu().numpy()
    lam, rho, iters, cg_tol, cg_max = 2.0, 5.0, 25, 1e-4, 25

    S = TVADMM(A, y, lam, rho, x0=x0, maxiter=iters, cg_tol=cg_tol, cg_maxiter=cg_max)
    counts, worst = [], {"x": 0.0, "trace": [0.0] * 13}
    for it in range(iters):
        xs, zs, us = (a.cpu().numpy().copy() for a in (S.x, S.z, S.u))
        S.step()
        k, tr_g, tol_g = S.cg_info["num_iter"], list(S.cg_trace), S.cg_tol_sq
        counts.append(k)
        # the solver's own decisions follow the reference's loop condition (scico/solver.py:393)
        assert len(tr_g) == k + 1 and all(t > tol_g for t in tr_g[:k]) and (k == cg_max or tr_g[k] <= tol_g)
        # the oracle's iteration from the same state, CG forced to the same depth
        xo, zo, uo, info = T.admm_tv_step(xs, zs, us, Ao, ATo, yn, lam, rho, cg_tol=0.0, cg_maxiter=k)
        assert info["num_iter"] == k
        # threshold: (f32(tol) * f32(||b||))^2 on both sides
        rhs = (ATo(yn) + f32(rho) * T.finite_difference_adj(zs - us)).astype(f32)
        tol_o = float((f32(cg_tol) * np.linalg.norm(rhs.ravel()).astype(f32)) ** 2)
        assert abs(tol_g - tol_o) <= 1e-5 * tol_o
        # <r, r> before each CG iteration: the same sequence at the start; deeper, the residual norm is the
        # quantity fp32 CG loses first (it sits at the rounding floor of A^T A p long before x does)
        for j in range(min(k, 12) + 1):
            dev = abs(info["trace"][j] / tr_g[j] - 1.0)
            worst["trace"][j] = max(worst["trace"][j], dev)
        # tolerance: the oracle's own sensitivity, from the same state at the same depth, to two rounding-level
        # changes the device also makes -- CG's inner products accumulated in fp64 instead of fp32, and the
        # forward projection's sums taken in another order (the C port's `fused` scatter) -- times 8, and never
        # below the 1e-4 of the fixed-depth tests: fp32 CG amplifies rounding by ~1.5 - 2.5x per iteration
        Af = lambda v: C.project_2d(v, Tb, A.ny, fused=True)  # noqa: E731
        x64 = _cg_fp64_sums(lambda v: (f32(rho) * T.finite_difference_adj(T.finite_difference(v)) + ATo(Af(v))).astype(f32),
                            rhs, xs, k)
        # ... and never below 1e-4 * 2^(depth - 10): the amplification alone (measured 1.5 - 2.5x per CG iteration;
        # three runs of this very test differ from each other by that much, the forward's atomics reorder sums)
        tol_x = min(max(1e-4 * 2.0 ** max(0, k - 10), 8.0 * O.rel_l2(x64, xo)), 2e-2)
        ex = O.rel_l2(S.x.cpu().numpy(), xo)
        worst["x"] = max(worst["x"], ex / tol_x)
        assert ex <= tol_x, (it, k, ex, tol_x)
        assert O.rel_l2(S.z.cpu().numpy(), zo) <= 10 * tol_x, (it, k)
        assert np.abs(S.u.cpu().numpy() - uo).max() <= 10 * tol_x * max(1.0, np.abs(uo).max()), (it, k)
    # the first inner products of every x-step (before fp32 CG has amplified anything) agree closely
    assert max(worst["trace"][:2]) <= 1e-2, worst
    assert any(c == 0 for c in counts) and any(c >= 10 for c in counts), counts  # the on/off pattern described above

    # free-running oracle on the same data: quality parity (two CPU arithmetic variants differ by 0.94 dB)
    x, z, u = T.admm_tv_init(x0n)
    for _ in range(iters):
        x, z, u, _info = T.admm_tv_step(x, z, u, Ao, ATo, yn, lam, rho, cg_tol=cg_tol, cg_maxiter=cg_max)
    snr = lambda ref, rec: 10 * np.log10(np.sum(ref ** 2) / np.sum((ref - rec) ** 2))  # noqa: E731
    s_fbp, s_gpu, s_cpu = snr(x_gt, x0n), snr(x_gt, np.clip(S.x.cpu().numpy(), 0, 1)), snr(x_gt, np.clip(x, 0, 1))
    # free-running results are chaotic at the 1e-2 level (docstring): gross-error bounds only.  Seen over repeated
    # runs: FBP 19.9 dB; GPU 32.4 - 33.3 dB; oracle 29.5 - 32.3 dB (fp32 vs fp64 inner products alone: 0.9 dB)
    assert s_gpu > s_fbp + 5.0 and s_cpu > s_fbp + 5.0, (s_fbp, s_gpu, s_cpu)
    assert abs(s_gpu - s_cpu) <= 6.0, (s_gpu, s_cpu)
    assert O.rel_l2(S.x.cpu().numpy(), x) <= 0.15


@pytest.mark.parametrize("dim,nonneg", [(3, False), (3, True), (2, False)])
def test_ladmm_iterations_match_oracle(cuda_device, dim, nonneg):
    import torch

    N, A, Ao, ATo, x_gt, y = _problem_2d() if dim == 2 else _problem_3d()
    mu, nu = TVLinearizedADMM.estimate_parameters(A, nu=1.0, factor=1.05, maxiter=40)
    lam, iters = 0.1, 25
    x0 = 0.1 * np.random.default_rng(5).standard_normal(N).astype(f32)
    S = TVLinearizedADMM(A, _t(torch, y, cuda_device), lam, mu, nu, x0=_t(torch, x0, cuda_device), nonneg=nonneg, maxiter=iters)
    x, z, u = T.ladmm_tv_init(x0, Ao)
    for _ in range(iters):
        x, z, u = T.ladmm_tv_step(x, z, u, Ao, ATo, y, lam, mu, nu, nonneg=nonneg)
    S.solve()
    assert O.rel_l2(S.x.cpu().numpy(), x) <= 1e-4
    assert O.rel_l2(S.z[0].cpu().numpy(), z[0]) <= 1e-4 and O.rel_l2(S.z[1].cpu().numpy(), z[1]) <= 1e-4
    assert O.rel_l2(S.u[0].cpu().numpy(), u[0]) <= 1e-4 and O.rel_l2(S.u[1].cpu().numpy(), u[1]) <= 1e-4


@pytest.mark.parametrize("dim,alpha", [(3, 1.0), (3, 10.0), (2, 4.0)])
def test_padmm_iterations_match_oracle(cuda_device, dim, alpha):
    """ct_3d_tv_padmm.py set-up (A = (C; alpha D), B = -I, zero start) for 25 iterations."""
    import torch

    N, A, Ao, ATo, x_gt, y = _problem_2d() if dim == 2 else _problem_3d()
    mu, nu = TVProximalADMM.estimate_parameters(A, alpha=alpha, factor=1.01, maxiter=40)
    assert abs(nu - 1.01) < 1e-12
    lam, rho, iters = 0.1, 0.05, 25
    S = TVProximalADMM(A, _t(torch, y, cuda_device), lam, rho, mu, nu, alpha=alpha, maxiter=iters, itstat=True)
    x, z, u, uo = T.padmm_tv_init(N, y.shape)
    for _ in range(iters):
        x, z, u, uo = T.padmm_tv_step(x, z, u, uo, Ao, ATo, y, lam, alpha, rho, mu, nu)
    S.solve()
    assert O.rel_l2(S.x.cpu().numpy(), x) <= 1e-4
    assert O.rel_l2(S.z[0].cpu().numpy(), z[0]) <= 1e-4 and O.rel_l2(S.z[1].cpu().numpy(), z[1]) <= 1e-4
    assert O.rel_l2(S.u[0].cpu().numpy(), u[0]) <= 1e-4 and O.rel_l2(S.u[1].cpu().numpy(), u[1]) <= 1e-4
    assert len(S.history) == iters and all(np.isfinite(h["objective"]) for h in S.history)


@pytest.mark.parametrize("solver,dim,fast", [("padmm", 3, True), ("padmm", 3, False), ("padmm", 2, True), ("ladmm", 3, False)])
def test_split_solver_fused_iteration_statistics_match_explicit_ones(cuda_device, solver, dim, fast):
    """Objective f(x) + g(z), primal residual ||A x + B z|| and dual residual (fast: ||z - z_old||, the reference's
    ProximalADMM default; otherwise ||C^T (z - z_old)||) as accumulated by the prox kernels (xct_*_prox_step_stat)
    against the same quantities evaluated with torch from the iterates (_padmm.py:148-177,294-345, _ladmm.py:130-213);
    the iterates are those of a run without statistics (up to the last bit: the two instantiations of a kernel need not
    contract the same products into FMAs)."""
    import torch

    from scico_b200.optimize import FiniteDifference

    N, A, Ao, ATo, x_gt, y = _problem_2d() if dim == 2 else _problem_3d()
    yt = _t(torch, y, cuda_device)
    lam, iters = 0.1, 12
    if solver == "padmm":
        alpha = 4.0
        mu, nu = TVProximalADMM.estimate_parameters(A, alpha=alpha, factor=1.01, maxiter=40)
        mk = lambda st: TVProximalADMM(A, yt, lam, 0.05, mu, nu, alpha=alpha, maxiter=iters, itstat=st,  # noqa: E731
                                       fast_dual_residual=fast)
    else:
        alpha = 1.0
        mu, nu = TVLinearizedADMM.estimate_parameters(A, nu=1.0, maxiter=40)
        mk = lambda st: TVLinearizedADMM(A, yt, lam, mu, nu, maxiter=iters, itstat=st)  # noqa: E731
    S, P = mk(True), mk(False)
    D = FiniteDifference(S.vol_shape)
    dn = lambda t: float(t.double().norm())  # noqa: E731
    for _ in range(iters):
        z0o, z1o = S.z0.clone(), S.z1.clone()
        S.step()
        P.step()
        h = S.history[-1]
        xv = S.x.reshape(S.vol_shape)
        pr = np.hypot(dn(A.project(S.x) - S.z0), dn(alpha * D(xv) - S.z1))
        obj = 0.5 * dn(S.z0 - yt) ** 2 + (lam / alpha) * float(torch.sqrt((S.z1.double() ** 2).sum(0)).sum())
        if fast:
            du = np.hypot(dn(S.z0 - z0o), dn(S.z1 - z1o))
        else:
            du = dn(A.back_project(S.z0 - z0o).reshape(S.vol_shape) + alpha * D.adj(S.z1 - z1o))
        assert abs(h["objective"] - obj) <= 1e-6 * abs(obj) + 1e-12
        assert abs(h["prml_rsdl"] - pr) <= 1e-5 * pr + 1e-9
        assert abs(h["dual_rsdl"] - du) <= 1e-5 * du + 1e-9
    for a, b in ((S.x, P.x), (S.z0, P.z0), (S.z1, P.z1), (S.u0, P.u0)):
        assert O.rel_l2(a.cpu().numpy(), b.cpu().numpy()) <= 1e-6


def test_padmm_estimate_matches_oracle_power_iteration(cuda_device):
    N, A, Ao, ATo, x_gt, y = _problem_3d()
    alpha = 7.0
    mu, _ = TVProximalADMM.estimate_parameters(A, alpha=alpha, factor=None, maxiter=60)
    v = np.random.default_rng(0).standard_normal(N).astype(f32)
    for _ in range(120):
        v /= np.linalg.norm(v)
        w = ATo(Ao(v)) + f32(alpha * alpha) * T.finite_difference_adj(T.finite_difference(v))
        ref = float(np.sum(v * w))
        v = w
    assert abs(mu - ref) <= 0.03 * ref
