"""Device-resident TV-regularised CT solvers around the projector pair:
:class:`TVPDHG`, :class:`TVADMM` (ADMM with a CG x-step), :class:`TVLinearizedADMM` and
:class:`TVProximalADMM`.

:class:`TVPDHG` restates ``scico.optimize.PDHG.step`` (``scico/optimize/_primaldual.py:219-231``) for
the problem every 3D CT example of the reference sets up,

    min_x  1/2 || A x - y ||^2  +  lam || D x ||_{2,1}   (+ optional x >= 0)

i.e. ``C = VerticalStack((A, D))``, ``D = FiniteDifference(input_shape, append=0)``
(``scico/linop/_diff.py:25-96``), ``g = Separable(SquaredL2Loss(y), lam * L21Norm())``
(``scico/loss.py:220-226``, ``scico/functional/_norm.py:254-263``), ``f = ZeroFunctional`` or
``NonNegativeIndicator``.  One iteration is one back projection, one forward projection and three
fused elementwise/stencil kernels (``scico_b200/csrc/xct_tv.cuh``); all state lives on the GPU and no
host synchronisation happens inside :meth:`step` (the reference syncs once per iteration statistic,
``scico/optimize/_common.py:305-312``; here statistics are optional and off by default).

With a :class:`~scico_b200.sharded.ViewShardedXRayTransform3D` (general matrices) the volume-shaped state is
z-slab sharded in the same way and the sinogram-shaped state is view-block sharded (first device run: next
round).  With a :class:`~scico_b200.sharded.SlabShardedXRayTransform3D` the state is z-slab sharded: the
projector needs no collective, the finite difference along axis 0 needs a one-plane halo per
iteration in each direction (``N1*N2*4`` bytes, point-to-point between neighbouring ranks).

The ADMM-family classes restate ``ADMM.step`` + ``LinearSubproblemSolver`` + ``scico.solver.cg``
(``scico/optimize/_admm.py:334-378``, ``_admmaux.py:206-269``, ``scico/solver.py:367-405``; the
problem of ``examples/scripts/ct_tv_admm.py``), ``LinearizedADMM.step`` (``_ladmm.py:253-277``) and
``ProximalADMM.step`` (``_padmm.py:349-363``; the problem of ``examples/scripts/ct_3d_tv_padmm.py``)
on the same fused kernels (``scico_b200/csrc/xct_solver.cuh``).  2D operators are handled as the
volume block ``(1, N0, N1)``: the axis-0 component of every gradient-shaped array stays exactly 0.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import numpy as np

from . import _lib

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class FiniteDifference:
    """``FiniteDifference(input_shape, append=0)`` on CUDA tensors (3D): ``(3, *shape) = D x`` and its
    adjoint, through ``xct_fd_forward`` / ``xct_fd_adjoint``."""

    def __init__(self, input_shape, is_first: bool = True, is_last: bool = True):
        self.input_shape = tuple(int(s) for s in input_shape)
        if len(self.input_shape) != 3:
            raise ValueError("scico_b200.optimize.FiniteDifference handles 3D volumes")
        self.output_shape = (3,) + self.input_shape
        self.blk = _lib.TvBlock(*self.input_shape, int(is_first), int(is_last))

    def __call__(self, x, hi_halo=None):
        out = torch.empty(self.output_shape, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().xct_fd_forward(ctypes.byref(self.blk), x.data_ptr(),
                                                 hi_halo.data_ptr() if hi_halo is not None else None,
                                                 out.data_ptr(), _stream(x.device)))
        return out

    def adj(self, z, lo_halo=None):
        out = torch.empty(self.input_shape, dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().xct_fd_adjoint(ctypes.byref(self.blk), z.data_ptr(),
                                                 lo_halo.data_ptr() if lo_halo is not None else None,
                                                 out.data_ptr(), _stream(z.device)))
        return out


class _TVSolver:
    """State and plumbing shared by the TV solvers: operator / sharding bookkeeping, halo exchange of
    boundary planes between neighbouring z-slabs, operator applications into preallocated buffers,
    global sums and the TV objective."""

    def _setup(self, A, y, x0=None):
        self.A = A
        self.sharded = hasattr(A, "slab")
        self.group = getattr(A, "group", None)
        self.rank = getattr(A, "rank", 0)
        self.world = getattr(A, "world_size", 1)
        in_shape = tuple(A.local_input_shape if self.sharded else A.input_shape)
        out_shape = tuple(A.local_output_shape if self.sharded else A.output_shape)
        if len(in_shape) not in (2, 3):
            raise ValueError(f"operator input must be 2D or 3D, got shape {in_shape}")
        if tuple(y.shape) != out_shape or not y.is_cuda:
            raise ValueError(f"y must be a CUDA tensor of shape {out_shape}")
        self.in_shape, self.out_shape = in_shape, out_shape
        self.vol_shape = in_shape if len(in_shape) == 3 else (1,) + in_shape  # 2D image = one slice
        self.dev = y.device
        self.y = y.to(torch.float32).contiguous()
        self.blk = _lib.TvBlock(*self.vol_shape, int(self.rank == 0), int(self.rank == self.world - 1))
        if x0 is None:
            self.x = torch.zeros(in_shape, dtype=torch.float32, device=self.dev)
        else:
            if tuple(x0.shape) != in_shape:
                raise ValueError(f"x0 must have shape {in_shape}")
            self.x = x0.to(device=self.dev, dtype=torch.float32).clone().contiguous()
        self.itnum = 0
        self._history = []
        self._stat_buf, self._stat_n, self._stat_pending = None, 0, []

    def _vol(self):
        return torch.empty(self.in_shape, dtype=torch.float32, device=self.dev)

    def _sino(self, zero=False):
        return (torch.zeros if zero else torch.empty)(self.out_shape, dtype=torch.float32, device=self.dev)

    def _grad(self):
        return torch.zeros((3,) + self.vol_shape, dtype=torch.float32, device=self.dev)

    def _grad_view(self, g):
        """Gradient-shaped state in the reference's shape: (3, N0, N1, N2) or (2, N0, N1)."""
        return g if len(self.in_shape) == 3 else g[1:, 0]

    # -- halo exchange (z-slab sharding only) ---------------------------------------------------
    def _neighbour(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def _exchange_planes(self, plane, send_to, recv_from):
        """Point-to-point exchange of one boundary plane: send ``plane`` to rank ``send_to`` (None: nobody),
        receive the matching plane of rank ``recv_from`` (None: nobody -> returns None).  NCCL moves device
        memory directly; under gloo (the CPU / one-GPU test rendezvous, which cannot send device pointers) the
        plane is staged through host memory."""
        staged = plane.is_cuda and dist.get_backend(self.group) == "gloo"
        src = plane.cpu() if staged else plane
        ops, recv = [], None
        if send_to is not None:
            ops.append(dist.P2POp(dist.isend, src, self._neighbour(send_to), self.group))
        if recv_from is not None:
            recv = torch.empty_like(src)
            ops.append(dist.P2POp(dist.irecv, recv, self._neighbour(recv_from), self.group))
        for req in dist.batch_isend_irecv(ops) if ops else []:
            req.wait()
        return recv.to(plane.device) if (staged and recv is not None) else recv

    def _halo_from_prev(self, plane_to_next):
        """Send ``plane_to_next`` to rank+1, receive the previous rank's plane (None on rank 0)."""
        if self.world == 1:
            return None
        return self._exchange_planes(plane_to_next, self.rank + 1 if self.rank + 1 < self.world else None,
                                     self.rank - 1 if self.rank > 0 else None)

    def _halo_from_next(self, plane_to_prev):
        if self.world == 1:
            return None
        return self._exchange_planes(plane_to_prev, self.rank - 1 if self.rank > 0 else None,
                                     self.rank + 1 if self.rank + 1 < self.world else None)

    def _lo_plane(self, g):
        """Halo for ``D^T g``: plane ``g[0][-1]`` of the previous slab."""
        return None if self.world == 1 else self._halo_from_prev(g[0, -1].contiguous())

    def _hi_plane(self, v):
        """Halo for ``D v``: plane ``v[0]`` of the next slab."""
        return None if self.world == 1 else self._halo_from_next(v[0].contiguous())

    @staticmethod
    def _ptr(t):
        return t.data_ptr() if t is not None else None

    # -- operator applications -------------------------------------------------------------------
    def _fwd(self, x, out):
        return self.A.project(x, out=out)  # plain and sharded operators write into the solver's own buffer

    def _adj(self, y, out):
        res = self.A.back_project(y, out=out)
        return res if res is not None else out

    def _sino_rows(self):
        """(inner, rows, lo, hi) of :c:func:`xct_l2_dual_step_stat`: the detector rows of the local sinogram block
        that count in global sums (z-slabs: the rows this rank owns; otherwise everything)."""
        if self.sharded and hasattr(self.A, "owned_rows") and len(self.out_shape) == 3:
            lo, hi = self.A.owned_rows
            return self.out_shape[2], self.out_shape[1], lo - self.A.rows[0], hi - self.A.rows[0]
        return self.y.numel(), 1, 0, 1

    # -- iteration statistics: device rows, read back on demand ------------------------------------
    _STAT_ROWS = 256

    @property
    def history(self):
        """List of per-iteration statistics (``iter``, ``objective``, ``prml_rsdl``, ``dual_rsdl``).  The sums behind
        them are accumulated on the device by the iteration's own kernels into one row per iteration; the rows are
        read back here (one copy for all pending iterations), so an iteration with statistics never waits for the
        host -- reading ``history`` inside a callback costs one synchronisation, as any display of the values would."""
        self._flush_stats()
        return self._history

    def _stat_row(self, width, to_entry):
        """Device pointer of a zeroed row of ``width`` doubles for this iteration's sums; ``to_entry(row values)``
        turns the row into the history entry when it is read back."""
        if self._stat_buf is None:
            self._stat_buf = torch.zeros((self._STAT_ROWS, width), dtype=torch.float64, device=self.dev)
        if self._stat_n == self._STAT_ROWS:
            self._flush_stats()
        row = self._stat_buf[self._stat_n]
        self._stat_n += 1
        self._stat_pending.append((self.itnum + 1, to_entry))
        return row

    def _flush_stats(self):
        if not self._stat_pending:
            return
        rows = self._stat_buf[: self._stat_n].cpu().tolist()  # stream-ordered after the kernels that wrote them
        for (it, to_entry), vals in zip(self._stat_pending, rows):
            entry = {"iter": it}
            entry.update(to_entry(vals))
            self._history.append(entry)
        self._stat_buf[: self._stat_n].zero_()
        self._stat_n, self._stat_pending = 0, []

    #: an iteration never needs the host (no scalar read back): it can be captured in a CUDA graph
    _graphable = False

    def solve(self, callback=None, use_graph: Optional[bool] = None):
        """Run ``maxiter`` iterations.

        ``use_graph``: capture ONE iteration in a CUDA graph and replay it (small problems are bound by
        launch and Python overhead, not by the kernels).  Possible when nothing in an iteration needs the
        host: single GPU, iteration statistics off, no callback, and not :class:`TVADMM`, whose CG stop
        test reads a scalar.  Default: off -- measured on a B200, plain stepping already keeps the GPU
        busy (2D 256^2 x 180 views PDHG: 12 000 iterations/s eager against 10 000 replayed), because no
        iteration waits for the host; the option exists for hosts with slower launch paths."""
        possible = (self._graphable and self.world == 1 and not self.sharded and not getattr(self, "itstat", False)
                    and callback is None)
        if use_graph is None:
            use_graph = False
        elif use_graph and not possible:
            raise ValueError("use_graph=True needs a single-GPU solver without iteration statistics or callback")
        done = 0
        if use_graph:
            self.step()  # warm-up outside the capture (plans, kernel attributes)
            done = 1
            graph = None
            try:
                torch.cuda.synchronize(self.dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self.step()  # recorded, not executed
                self.itnum -= 1
            except RuntimeError:  # capture refused: plain launches
                graph = None
            if graph is not None:
                for _ in range(self.maxiter - done):
                    graph.replay()
                    self.itnum += 1
                return self.x
        for _ in range(self.maxiter - done):
            self.step()
            if callback is not None:
                callback(self)
        return self.x

    # -- statistics (host syncs) -----------------------------------------------------------------
    def _sum(self, t) -> float:
        s = t.double().sum()
        if self.world > 1:
            dist.all_reduce(s, group=self.group)
        return float(s.item())

    def _owned(self, t):
        """Sinogram-shaped array restricted to the detector rows this rank owns (shared rows are
        counted once in global sums)."""
        if self.sharded and hasattr(self.A, "owned_rows") and tuple(t.shape) == self.out_shape:
            lo, hi = self.A.owned_rows  # z-slabs; a view block is owned entirely (views are disjoint)
            return t[:, lo - self.A.rows[0]: hi - self.A.rows[0]]
        return t

    def _norm(self, t) -> float:
        return math.sqrt(self._sum(self._owned(t).double() ** 2))

    def _fd(self):
        return FiniteDifference(self.vol_shape, self.rank == 0, self.rank == self.world - 1)

    def _l21(self, g) -> float:
        return self._sum(torch.sqrt((g.double() ** 2).sum(dim=0)))

    def objective(self, x=None) -> float:
        """``1/2||Ax - y||^2 + lam ||Dx||_{2,1}`` at ``x`` (default: the current iterate)."""
        x = self.x if x is None else x
        r = self._owned(self.A.project(x) - self.y)
        xv = x.reshape(self.vol_shape)
        d = self._fd()(xv, self._hi_plane(xv))
        return 0.5 * self._sum(r.double() ** 2) + self.lam * self._l21(d)

    @staticmethod
    def operator_norm_sq(A, dscale: float = 1.0, maxiter: int = 100, seed: int = 0) -> float:
        """``|| (A; dscale*D) ||_2^2`` by power iteration of ``A^T A + dscale^2 D^T D`` on the device
        (``scico/linop/_util.py:27-110``; 100 iterations by default like the reference).  Power iteration
        approaches the norm from BELOW: fewer iterations under-estimate it, which makes every step size
        derived from it too large -- pass a smaller ``maxiter`` only together with a safety margin."""
        h = _TVSolver()
        sharded = hasattr(A, "slab")
        shape = tuple(A.local_input_shape if sharded else A.input_shape)
        vshape = shape if len(shape) == 3 else (1,) + shape
        h.world, h.rank, h.group = getattr(A, "world_size", 1), getattr(A, "rank", 0), getattr(A, "group", None)
        dev = torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator(device=dev).manual_seed(seed + h.rank)
        v = torch.randn(shape, device=dev, generator=g)
        D = FiniteDifference(vshape, h.rank == 0, h.rank == h.world - 1)
        mu = 1.0
        for _ in range(maxiter):
            v = v / math.sqrt(h._sum(v * v))
            vv = v.reshape(vshape)
            dz = D(vv, h._hi_plane(vv))
            w = A.back_project(A.project(v)) + (dscale * dscale) * D.adj(dz, h._lo_plane(dz)).reshape(shape)
            mu = h._sum(v * w)
            v = w
        return mu


class TVPDHG(_TVSolver):
    """PDHG for ``1/2||Ax - y||^2 + lam ||Dx||_{2,1}`` with the state on the GPU.

    Args:
        A: ``XRayTransform3D`` / ``XRayTransform2D`` (single GPU) or ``SlabShardedXRayTransform3D``
            (one slab per rank).
        y: measured sinogram, CUDA tensor of ``A``'s (local) output shape.
        lam: TV weight.  tau, sigma: step sizes (``tau * sigma * ||C||^2 < 1``; see
            :meth:`estimate_parameters`).  alpha: relaxation (reference default 1.0).
        nonneg: use ``f = NonNegativeIndicator`` instead of ``ZeroFunctional``.
        x0: initial volume (default zeros).  maxiter: iterations run by :meth:`solve`.
        itstat: record objective / residual norms every iteration like the reference's ``itstat_options`` (accumulated
            by the iteration's own kernels; ``A x`` follows from ``A xbar`` by linearity, so no extra projection; one
            host read per iteration).
    """

    _graphable = True

    def __init__(self, A, y, lam: float, tau: float, sigma: float, alpha: float = 1.0, nonneg: bool = False,
                 x0=None, maxiter: int = 100, itstat: bool = False):
        self._setup(A, y, x0)
        self.lam, self.tau, self.sigma, self.alpha = float(lam), float(tau), float(sigma), float(alpha)
        self.nonneg = bool(nonneg)
        self.maxiter = int(maxiter)
        self.itstat = bool(itstat)
        self.xbar = self.x.clone()
        self.z0 = self._sino(zero=True)
        self.z1 = self._grad()
        self.atz = self._vol()
        self.ax = self._sino()
        if self.itstat:
            # A x of the current iterate, kept up to date by xct_l2_dual_step_stat from A xbar (no second forward
            # projection per iteration); five device doubles: ||dx||^2, ||dz1||^2, ||dz0||^2, ||Ax - y||^2, ||Dx||_{2,1}
            self.ax_x = self._sino(zero=True) if x0 is None else self.A.project(self.x).clone()

    def step(self):
        """One PDHG iteration (``_primaldual.py:219-231``).  With ``itstat`` the same kernels also accumulate the
        reference's iteration statistics (``_primaldual.py:175-217``) on the device: no copies of the old iterates,
        no extra projection, no host synchronisation (the sums are read back when ``history`` is looked at)."""
        L = _lib.lib()
        blk = ctypes.byref(self.blk)
        with torch.cuda.device(self.dev):
            st = _stream(self.dev)
            if self.itstat:
                tau, sigma, lam = self.tau, self.sigma, self.lam
                row = self._stat_row(5, lambda v: {  # ||dx||^2, ||dz1||^2, ||dz0||^2, ||Ax - y||^2, ||Dx||_{2,1}
                    "objective": 0.5 * v[3] + lam * v[4], "prml_rsdl": math.sqrt(v[0]) / tau,
                    "dual_rsdl": math.sqrt(v[2] + v[1]) / sigma})
                sp = lambda i: row.data_ptr() + 8 * i  # noqa: E731
            self.atz = self._adj(self.z0, self.atz)                       # A^T z0
            lo = self._lo_plane(self.z1)                                  # z1[0][-1] of the previous slab
            if self.itstat:
                _lib.check(L.xct_tv_primal_step_stat(blk, self.x.data_ptr(), self.xbar.data_ptr(), self.atz.data_ptr(),
                                                     self.z1.data_ptr(), self._ptr(lo), self.tau, self.alpha,
                                                     int(self.nonneg), sp(0), st))
            else:
                _lib.check(L.xct_tv_primal_step(blk, self.x.data_ptr(), self.xbar.data_ptr(), self.atz.data_ptr(),
                                                self.z1.data_ptr(), self._ptr(lo), self.tau, self.alpha,
                                                int(self.nonneg), st))
            self.ax = self._fwd(self.xbar, self.ax)                       # A xbar
            hi = self._hi_plane(self.xbar.reshape(self.vol_shape))        # xbar[n0] of the next slab
            if self.itstat:
                _lib.check(L.xct_tv_dual_step_stat(blk, self.z1.data_ptr(), self.xbar.data_ptr(), self._ptr(hi),
                                                   self.sigma, self.lam, sp(1), st))
                _lib.check(L.xct_l2_dual_step_stat(self.z0.numel(), self.z0.data_ptr(), self.ax.data_ptr(),
                                                   self.y.data_ptr(), self.sigma, self.ax_x.data_ptr(), self.alpha,
                                                   *self._sino_rows(), sp(2), st))
                xv = self.x.reshape(self.vol_shape)
                _lib.check(L.xct_tv_norm(blk, self.x.data_ptr(), self._ptr(self._hi_plane(xv)), sp(4), st))
            else:
                _lib.check(L.xct_tv_dual_step(blk, self.z1.data_ptr(), self.xbar.data_ptr(), self._ptr(hi),
                                              self.sigma, self.lam, st))
                _lib.check(L.xct_l2_dual_step(self.z0.numel(), self.z0.data_ptr(), self.ax.data_ptr(),
                                              self.y.data_ptr(), self.sigma, st))
            if self.itstat and self.world > 1:
                dist.all_reduce(row, group=self.group)
        self.itnum += 1

    @staticmethod
    def estimate_parameters(A, ratio: float = 1.0, factor: Optional[float] = 1.01, maxiter: int = 100, seed: int = 0):
        """(tau, sigma) from ``||C||_2`` by power iteration of ``C^T C = A^T A + D^T D`` on the device
        (``_primaldual.py:234-288``, ``scico/linop/_util.py:27-110``; 100 iterations like the reference:
        the estimate converges from below, so a truncated iteration makes ``tau * sigma * ||C||^2`` larger,
        not smaller)."""
        cnorm = math.sqrt(_TVSolver.operator_norm_sq(A, 1.0, maxiter, seed))
        factor = 1.0 if factor is None else factor
        # reference formula (_primaldual.py:286-288); note that factor > 1 loosens tau*sigma*||C||^2 < 1,
        # pass factor < 1 for a strict bound
        tau = math.sqrt(factor / ratio) / cnorm
        return tau, ratio * tau


class TVADMM(_TVSolver):
    """ADMM with a conjugate-gradient x-step for ``1/2||Ax - y||^2 + lam ||Dx||_{2,1}``
    (``examples/scripts/ct_tv_admm.py:68-84``: ``f = SquaredL2Loss(y, A)``, ``g_list = [lam*L21Norm()]``,
    ``C_list = [FiniteDifference(append=0)]``, ``rho_list = [rho]``, ``LinearSubproblemSolver``).

    x-step: CG on ``(A^T A + rho D^T D) x = A^T y + rho D^T (z - u)`` started from the current ``x``
    (``_admmaux.py:231-269``, ``scico/solver.py:367-405``); ``A^T y`` is computed once (the reference
    recomputes it every iteration, ``_admmaux.py:246-248``).  One CG iteration is one forward + one back
    projection and three fused kernels; ``alpha``/``beta`` stay on the device, the host reads back only
    the squared residual norm that decides termination (the reference's ``while`` test does the same).

    Args:
        A, y, lam, x0, maxiter, itstat: as :class:`TVPDHG` (``x0`` e.g. ``clip(A.fbp(y), 0, 1)``).
        rho: ADMM penalty parameter.  cg_tol, cg_maxiter: ``cg_kwargs`` of the reference
            (defaults of ``LinearSubproblemSolver``: 1e-4, 100).
    """

    def __init__(self, A, y, lam: float, rho: float, x0=None, maxiter: int = 100, cg_tol: float = 1e-4,
                 cg_maxiter: int = 100, itstat: bool = False):
        self._setup(A, y, x0)
        self.lam, self.rho = float(lam), float(rho)
        self.maxiter, self.cg_tol, self.cg_maxiter = int(maxiter), float(cg_tol), int(cg_maxiter)
        self.itstat = bool(itstat)
        xv = self.x.reshape(self.vol_shape)
        self.z1 = self._fd()(xv, self._hi_plane(xv))  # z_init: z = C x0 (_admm.py:297-316)
        self.u1 = self._grad()                        # u_init: zeros (_admm.py:318-332)
        self.aty = self.A.back_project(self.y).contiguous()
        self.rhs, self.r, self.p, self.q, self.atq = (self._vol() for _ in range(5))
        self.ax = self._sino()
        self.scal = torch.zeros(8, dtype=torch.float64, device=self.dev)  # CG inner products (device)
        self.cg_info = {"num_iter": 0, "rel_res": 0.0}
        self.cg_iters_total = 0
        self.cg_trace, self.cg_tol_sq = [], 0.0
        self.ax_x = None

    def _allreduce(self, i):
        if self.world > 1:
            dist.all_reduce(self.scal[i:i + 1], group=self.group)

    def _halos(self, v):
        """(lo, hi) halo planes of ``v`` for ``D^T D v``: ``v[-1]`` of the previous, ``v[0]`` of the next slab."""
        if self.world == 1:
            return None, None
        vv = v.reshape(self.vol_shape)
        return self._halo_from_prev(vv[-1].contiguous()), self._halo_from_next(vv[0].contiguous())

    def _xstep(self):
        L, blk, st, n = _lib.lib(), ctypes.byref(self.blk), _stream(self.dev), self.x.numel()
        sc = self.scal
        sp = lambda i: sc.data_ptr() + 8 * i  # noqa: E731  (slots 0-2: r.r ring, 3-4: p.q ring, 5: ||b||^2)
        sc.zero_()
        zu_lo = None if self.world == 1 else self._halo_from_prev((self.z1[0, -1] - self.u1[0, -1]).contiguous())
        _lib.check(L.xct_admm_rhs(blk, self.aty.data_ptr(), self.z1.data_ptr(), self.u1.data_ptr(), self._ptr(zu_lo),
                                  self.rho, self.rhs.data_ptr(), sp(5), st))
        self.ax = self._fwd(self.x, self.ax)
        if self.itstat:
            # A x for the objective of the iteration statistics: anchored here on the exact A x_k, then carried
            # through the CG updates x += alpha p with the A p every CG iteration computes anyway
            # (A x += alpha A p) -- no forward projection of its own
            if self.ax_x is None:
                self.ax_x = torch.empty_like(self.ax)
            self.ax_x.copy_(self.ax)
        self.atq = self._adj(self.ax, self.atq)
        lo, hi = self._halos(self.x)
        _lib.check(L.xct_cg_init(blk, self.x.data_ptr(), self._ptr(lo), self._ptr(hi), self.atq.data_ptr(),
                                 self.rhs.data_ptr(), self.rho, self.r.data_ptr(), self.p.data_ptr(), sp(0), st))
        self._allreduce(0)
        self._allreduce(5)
        host = sc.cpu()
        num, bn2 = float(host[0]), float(host[5])
        # the reference's stop test in its own fp32 arithmetic (scico/solver.py:384-391): bn = ||b|| and
        # num = <r, r> are float32 scalars there, termination_tol_sq = maximum(tol * bn, atol) ** 2; the
        # device accumulates both sums in fp64 and they are rounded to fp32 here, before the comparison
        f32 = np.float32
        tol_sq = float((f32(self.cg_tol) * f32(math.sqrt(bn2))) ** 2)
        num = float(f32(num))
        self.cg_trace = [num]  # <r, r> before every CG iteration of this x-step (fp32-rounded)
        self.cg_tol_sq = tol_sq
        k = 0
        while k < self.cg_maxiter and num > tol_sq:
            cur, nxt, pq, pq_nxt = k % 3, (k + 1) % 3, 3 + k % 2, 3 + (k + 1) % 2
            self.ax = self._fwd(self.p, self.ax)
            self.atq = self._adj(self.ax, self.atq)
            lo, hi = self._halos(self.p)
            _lib.check(L.xct_cg_lhs(blk, self.p.data_ptr(), self._ptr(lo), self._ptr(hi), self.atq.data_ptr(),
                                    self.rho, self.q.data_ptr(), sp(pq), sp(nxt), st))
            self._allreduce(pq)
            _lib.check(L.xct_cg_update_xr(n, self.x.data_ptr(), self.r.data_ptr(), self.p.data_ptr(), self.q.data_ptr(),
                                          sp(cur), sp(pq), sp(nxt), sp(pq_nxt), st))
            if self.itstat:
                self.ax_x.addcmul_(self.ax, (sc[cur] / sc[pq]).to(torch.float32))  # alpha of this CG iteration, on the device
            self._allreduce(nxt)
            _lib.check(L.xct_cg_update_p(n, self.p.data_ptr(), self.r.data_ptr(), sp(cur), sp(nxt), st))
            num = float(f32(sc[nxt].item()))
            self.cg_trace.append(num)
            k += 1
        self.cg_info = {"num_iter": k, "rel_res": math.sqrt(num / bn2) if bn2 > 0 else 0.0}
        self.cg_iters_total += k

    def step(self):
        """One ADMM iteration (``_admm.py:334-378`` with ``alpha = 1``)."""
        if self.itstat:
            z_old = self.z1.clone()
        with torch.cuda.device(self.dev):
            self._xstep()
            xv = self.x.reshape(self.vol_shape)
            hi = self._hi_plane(xv)
            _lib.check(_lib.lib().xct_grad_prox_step(ctypes.byref(self.blk), self.x.data_ptr(), self._ptr(hi),
                                                     self.z1.data_ptr(), self.u1.data_ptr(), None, 1.0,
                                                     self.lam / self.rho, 1.0, _lib.SPLIT_ADMM, _stream(self.dev)))
        self.itnum += 1
        if self.itstat:
            D = self._fd()
            xv = self.x.reshape(self.vol_shape)
            pr = math.sqrt(self.rho) * self._norm(D(xv, self._hi_plane(xv)) - self.z1)  # _admm.py:253-277
            dz = self.z1 - z_old
            du = self.rho * self._norm(D.adj(dz, self._lo_plane(dz)))                     # _admm.py:279-295
            r = self._owned(self.ax_x - self.y)                                           # ax_x = A x, see _xstep
            obj = 0.5 * self._sum(r.double() ** 2) + self.lam * self._l21(self.z1)        # f(x) + g(z), _admm.py:215-251
            self.history.append({"iter": self.itnum, "objective": obj, "prml_rsdl": pr, "dual_rsdl": du,
                                 "cg_iters": self.cg_info["num_iter"], "cg_rel_res": self.cg_info["rel_res"]})

    @property
    def z(self):
        return self._grad_view(self.z1)

    @property
    def u(self):
        return self._grad_view(self.u1)


class _TVSplitSolver(_TVSolver):
    """Shared by :class:`TVLinearizedADMM` and :class:`TVProximalADMM`: the split variable is the pair
    (sinogram block, gradient block) of ``(A; dscale*D) x``; ``w0`` / ``w1`` hold the array the next
    x-step applies the adjoint to.  One iteration = one back projection, one forward projection and
    three fused kernels, no host synchronisation."""

    _graphable = True

    def _alloc(self):
        self.z0, self.u0, self.w0 = (self._sino(zero=True) for _ in range(3))
        self.z1, self.u1, self.w1 = (self._grad() for _ in range(3))
        self.atq, self.ax = self._vol(), self._sino()

    #: dual residual as ``||z - z_old||`` (the reference's ``fast_dual_residual``, ProximalADMM's default) instead of
    #: ``||C^T (z - z_old)||``, which costs copies of the old ``z`` and one more back projection per iteration
    fast_dual_residual = False

    def _iterate(self, mode, step, dscale, thr, c, inv_nu):
        L, blk = _lib.lib(), ctypes.byref(self.blk)
        slow_dual = self.itstat and not self.fast_dual_residual
        if slow_dual:
            z_old = (self.z0.clone(), self.z1.clone())
        with torch.cuda.device(self.dev):
            st = _stream(self.dev)
            self.atq = self._adj(self.w0, self.atq)
            lo = self._lo_plane(self.w1)
            _lib.check(L.xct_grad_primal_step(blk, self.x.data_ptr(), self.atq.data_ptr(), self.w1.data_ptr(),
                                              self._ptr(lo), step, dscale, int(self.nonneg), st))
            self.ax = self._fwd(self.x, self.ax)
            hi = self._hi_plane(self.x.reshape(self.vol_shape))
            if self.itstat:
                # the statistics' sums come out of the two prox kernels: (||Cx - z||^2, ||dz||^2, g(z) sum) of the
                # gradient block, then of the sinogram block; read back when `history` is looked at
                lam_g = self.lam / dscale
                if slow_dual:
                    row = self._stat_row(7, lambda v: {"objective": 0.5 * v[5] + lam_g * v[2],
                                                       "prml_rsdl": math.sqrt(v[3] + v[0]), "dual_rsdl": math.sqrt(v[6])})
                else:
                    row = self._stat_row(7, lambda v: {"objective": 0.5 * v[5] + lam_g * v[2],
                                                       "prml_rsdl": math.sqrt(v[3] + v[0]), "dual_rsdl": math.sqrt(v[4] + v[1])})
                sp = row.data_ptr()
                _lib.check(L.xct_grad_prox_step_stat(blk, self.x.data_ptr(), self._ptr(hi), self.z1.data_ptr(),
                                                     self.u1.data_ptr(), self.w1.data_ptr(), dscale, thr, inv_nu, mode, sp, st))
                _lib.check(L.xct_sino_prox_step_stat(self.z0.numel(), self.ax.data_ptr(), self.y.data_ptr(), self.z0.data_ptr(),
                                                     self.u0.data_ptr(), self.w0.data_ptr(), c, inv_nu, mode,
                                                     *self._sino_rows(), sp + 24, st))
            else:
                _lib.check(L.xct_grad_prox_step(blk, self.x.data_ptr(), self._ptr(hi), self.z1.data_ptr(), self.u1.data_ptr(),
                                                self.w1.data_ptr(), dscale, thr, inv_nu, mode, st))
                _lib.check(L.xct_sino_prox_step(self.z0.numel(), self.ax.data_ptr(), self.y.data_ptr(), self.z0.data_ptr(),
                                                self.u0.data_ptr(), self.w0.data_ptr(), c, inv_nu, mode, st))
            if self.itstat:
                if slow_dual:  # ||C^T (z - z_old)||^2 (this rank's voxels) into the row's last slot
                    e0, e1 = self.z0 - z_old[0], self.z1 - z_old[1]
                    ct = self.A.back_project(e0).reshape(self.vol_shape) + dscale * self._fd().adj(e1, self._lo_plane(e1))
                    row[6] = (ct.double() ** 2).sum()
                if self.world > 1:
                    dist.all_reduce(row, group=self.group)
        self.itnum += 1

    @property
    def z(self):
        return self.z0, self._grad_view(self.z1)

    @property
    def u(self):
        return self.u0, self._grad_view(self.u1)


class TVLinearizedADMM(_TVSplitSolver):
    """Linearized ADMM (``scico/optimize/_ladmm.py:253-277``) for ``C = VerticalStack((A, D))``,
    ``g = Separable(SquaredL2Loss(y), lam*L21Norm())``, ``f = ZeroFunctional`` or ``NonNegativeIndicator``.
    ``mu / nu`` must be below ``1 / ||C||^2`` (see :meth:`estimate_parameters`).  ``C x`` of the previous
    iteration is reused, so one iteration is one forward and one back projection (the reference applies
    ``C`` twice)."""

    def __init__(self, A, y, lam: float, mu: float, nu: float, x0=None, nonneg: bool = False, maxiter: int = 100,
                 itstat: bool = False):
        self._setup(A, y, x0)
        self.lam, self.mu, self.nu = float(lam), float(mu), float(nu)
        self.nonneg, self.maxiter, self.itstat = bool(nonneg), int(maxiter), bool(itstat)
        self._alloc()
        # z_init: z = C x0, u = 0 (_ladmm.py:216-251)  =>  w = (C x0 - z) + u = 0 exactly
        self.z0.copy_(self.A.project(self.x))
        xv = self.x.reshape(self.vol_shape)
        self.z1.copy_(self._fd()(xv, self._hi_plane(xv)))

    def step(self):
        self._iterate(_lib.SPLIT_LADMM, self.mu / self.nu, 1.0, self.lam * self.nu, self.nu, 1.0)

    @staticmethod
    def estimate_parameters(A, nu: float = 1.0, factor: float = 1.01, maxiter: int = 100, seed: int = 0):
        """(mu, nu) with ``mu = nu / (factor ||C||^2)`` (``maxiter``: see :meth:`_TVSolver.operator_norm_sq`)."""
        return nu / (factor * _TVSolver.operator_norm_sq(A, 1.0, maxiter, seed)), nu


class TVProximalADMM(_TVSplitSolver):
    """Proximal ADMM (``scico/optimize/_padmm.py:349-363``) for the splitting of
    ``examples/scripts/ct_3d_tv_padmm.py:96-120``: ``A_stack = VerticalStack((A, alpha*D))``, ``B = -I``,
    ``c = 0``, ``f = ZeroFunctional`` (or ``NonNegativeIndicator``),
    ``g = Separable(SquaredL2Loss(y), (lam/alpha)*L21Norm())``; ``x, z, u`` start at zero unless ``x0``
    is given (``_padmm.py:106-118``)."""

    def __init__(self, A, y, lam: float, rho: float, mu: float, nu: float, alpha: float = 1.0, x0=None,
                 nonneg: bool = False, maxiter: int = 100, itstat: bool = False, fast_dual_residual: bool = True):
        self._setup(A, y, x0)
        self.lam, self.rho, self.mu, self.nu, self.alpha = float(lam), float(rho), float(mu), float(nu), float(alpha)
        self.nonneg, self.maxiter, self.itstat = bool(nonneg), int(maxiter), bool(itstat)
        self.fast_dual_residual = bool(fast_dual_residual)  # the reference's default (_padmm.py:70,330-345)
        self._alloc()

    def step(self):
        plam = 1.0 / (self.rho * self.nu)
        self._iterate(_lib.SPLIT_PADMM, 1.0 / self.mu, self.alpha, (self.lam / self.alpha) * plam, plam, 1.0 / self.nu)

    @staticmethod
    def estimate_parameters(A, alpha: float = 1.0, factor: Optional[float] = 1.01, maxiter: int = 100, seed: int = 0):
        """(mu, nu) = factor * (``||(A; alpha D)||^2``, ``||-I||^2 = 1``) (``_padmm.py:365-412``)."""
        mu = _TVSolver.operator_norm_sq(A, alpha, maxiter, seed)
        f = 1.0 if factor is None else factor
        return f * mu, f * 1.0
