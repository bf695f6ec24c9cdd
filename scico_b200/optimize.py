"""Device-resident TV-regularised PDHG for CT around the projector pair.

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

With a :class:`~scico_b200.sharded.SlabShardedXRayTransform3D` the state is z-slab sharded: the
projector needs no collective, the finite difference along axis 0 needs a one-plane halo per
iteration in each direction (``N1*N2*4`` bytes, point-to-point between neighbouring ranks).
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


class TVPDHG:
    """PDHG for ``1/2||Ax - y||^2 + lam ||Dx||_{2,1}`` with the state on the GPU.

    Args:
        A: ``XRayTransform3D`` (single GPU) or ``SlabShardedXRayTransform3D`` (one slab per rank).
        y: measured sinogram, CUDA tensor of ``A``'s (local) output shape.
        lam: TV weight.  tau, sigma: step sizes (``tau * sigma * ||C||^2 < 1``; see
            :meth:`estimate_parameters`).  alpha: relaxation (reference default 1.0).
        nonneg: use ``f = NonNegativeIndicator`` instead of ``ZeroFunctional``.
        x0: initial volume (default zeros).  maxiter: iterations run by :meth:`solve`.
        itstat: record objective / residual norms every iteration (host sync, extra forward).
    """

    def __init__(self, A, y, lam: float, tau: float, sigma: float, alpha: float = 1.0, nonneg: bool = False,
                 x0=None, maxiter: int = 100, itstat: bool = False):
        self.A = A
        self.sharded = hasattr(A, "slab")
        self.group = getattr(A, "group", None)
        self.rank = getattr(A, "rank", 0)
        self.world = getattr(A, "world_size", 1)
        in_shape = A.local_input_shape if self.sharded else A.input_shape
        out_shape = A.local_output_shape if self.sharded else A.output_shape
        if tuple(y.shape) != tuple(out_shape) or not y.is_cuda:
            raise ValueError(f"y must be a CUDA tensor of shape {tuple(out_shape)}")
        self.dev = y.device
        self.y = y.to(torch.float32).contiguous()
        self.lam, self.tau, self.sigma, self.alpha = float(lam), float(tau), float(sigma), float(alpha)
        self.nonneg = bool(nonneg)
        self.maxiter = int(maxiter)
        self.itstat = bool(itstat)
        self.blk = _lib.TvBlock(*in_shape, int(self.rank == 0), int(self.rank == self.world - 1))
        self.x = torch.zeros(in_shape, dtype=torch.float32, device=self.dev) if x0 is None \
            else x0.to(torch.float32).clone().contiguous()
        self.xbar = self.x.clone()
        self.z0 = torch.zeros(out_shape, dtype=torch.float32, device=self.dev)
        self.z1 = torch.zeros((3,) + tuple(in_shape), dtype=torch.float32, device=self.dev)
        self.atz = torch.empty(in_shape, dtype=torch.float32, device=self.dev)
        self.ax = torch.empty(out_shape, dtype=torch.float32, device=self.dev)
        self.x_old = None
        self.itnum = 0
        self.history = []
        self._dx2 = None

    # -- halo exchange (z-slab sharding only) ---------------------------------------------------
    def _neighbour(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def _halo_from_prev(self, plane_to_next):
        """Send ``plane_to_next`` to rank+1, receive the previous rank's plane (None on rank 0)."""
        if self.world == 1:
            return None
        ops, recv = [], None
        if self.rank + 1 < self.world:
            ops.append(dist.P2POp(dist.isend, plane_to_next, self._neighbour(self.rank + 1), self.group))
        if self.rank > 0:
            recv = torch.empty_like(plane_to_next)
            ops.append(dist.P2POp(dist.irecv, recv, self._neighbour(self.rank - 1), self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return recv

    def _halo_from_next(self, plane_to_prev):
        if self.world == 1:
            return None
        ops, recv = [], None
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, plane_to_prev, self._neighbour(self.rank - 1), self.group))
        if self.rank + 1 < self.world:
            recv = torch.empty_like(plane_to_prev)
            ops.append(dist.P2POp(dist.irecv, recv, self._neighbour(self.rank + 1), self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return recv

    # -- operator applications -------------------------------------------------------------------
    def _fwd(self, x, out):
        return self.A.project(x) if self.sharded else self.A.project(x, out=out)

    def _adj(self, y, out):
        return self.A.back_project(y) if self.sharded else self.A.back_project(y, out=out)

    def step(self):
        """One PDHG iteration (``_primaldual.py:219-231``)."""
        L = _lib.lib()
        if self.itstat:
            self.x_old = self.x.clone()
            z0_old, z1_old = self.z0.clone(), self.z1.clone()
        with torch.cuda.device(self.dev):
            st = _stream(self.dev)
            self.atz = self._adj(self.z0, self.atz)                       # A^T z0
            lo = self._halo_from_prev(self.z1[0, -1].contiguous())        # z1[0][-1] of the previous slab
            _lib.check(L.xct_tv_primal_step(ctypes.byref(self.blk), self.x.data_ptr(), self.xbar.data_ptr(),
                                            self.atz.data_ptr(), self.z1.data_ptr(),
                                            lo.data_ptr() if lo is not None else None,
                                            self.tau, self.alpha, int(self.nonneg), st))
            self.ax = self._fwd(self.xbar, self.ax)                       # A xbar
            hi = self._halo_from_next(self.xbar[0].contiguous())          # xbar[n0] of the next slab
            _lib.check(L.xct_tv_dual_step(ctypes.byref(self.blk), self.z1.data_ptr(), self.xbar.data_ptr(),
                                          hi.data_ptr() if hi is not None else None, self.sigma, self.lam, st))
            _lib.check(L.xct_l2_dual_step(self.z0.numel(), self.z0.data_ptr(), self.ax.data_ptr(),
                                          self.y.data_ptr(), self.sigma, st))
        self.itnum += 1
        if self.itstat:
            pr = self._norm(self.x - self.x_old) / self.tau
            du = math.sqrt(self._norm(self.z0 - z0_old) ** 2 + self._norm(self.z1 - z1_old) ** 2) / self.sigma
            self.history.append({"iter": self.itnum, "objective": self.objective(), "prml_rsdl": pr, "dual_rsdl": du})

    def solve(self, callback=None):
        for _ in range(self.maxiter):
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

    def _norm(self, t) -> float:
        if self.sharded and t.shape == self.z0.shape:  # count shared detector rows once
            lo, hi = self.A.owned_rows
            t = t[:, lo - self.A.rows[0]: hi - self.A.rows[0]]
        return math.sqrt(self._sum(t.double() ** 2))

    def objective(self, x=None) -> float:
        """``f(x) + g(Cx)`` = ``1/2||Ax - y||^2 + lam ||Dx||_{2,1}`` (``_primaldual.py:172-189``)."""
        x = self.x if x is None else x
        r = (self.A.project(x) - self.y)
        if self.sharded:
            lo, hi = self.A.owned_rows
            r = r[:, lo - self.A.rows[0]: hi - self.A.rows[0]]
        hi_h = self._halo_from_next(x[0].contiguous())
        d = FiniteDifference(x.shape, self.rank == 0, self.rank == self.world - 1)(x, hi_h)
        return 0.5 * self._sum(r.double() ** 2) + self.lam * self._sum(torch.sqrt((d.double() ** 2).sum(dim=0)))

    @staticmethod
    def estimate_parameters(A, ratio: float = 1.0, factor: Optional[float] = 1.01, maxiter: int = 20, seed: int = 0):
        """(tau, sigma) from ``||C||_2`` by power iteration of ``C^T C = A^T A + D^T D`` on the device
        (``_primaldual.py:234-288``, ``scico/linop/_util.py:27-110``; the reference runs 100
        iterations, 20 are enough for the 1 % safety factor)."""
        sharded = hasattr(A, "slab")
        shape = A.local_input_shape if sharded else A.input_shape
        rank, world = getattr(A, "rank", 0), getattr(A, "world_size", 1)
        group = getattr(A, "group", None)
        dev = torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator(device=dev).manual_seed(seed + rank)
        v = torch.randn(shape, device=dev, generator=g)
        D = FiniteDifference(shape, rank == 0, rank == world - 1)
        helper = TVPDHG.__new__(TVPDHG)
        helper.world, helper.rank, helper.group = world, rank, group

        def gsum(t):
            s = t.double().sum()
            if world > 1:
                dist.all_reduce(s, group=group)
            return float(s.item())

        factor = 1.0 if factor is None else factor
        mu = 1.0
        for _ in range(maxiter):
            v = v / math.sqrt(gsum(v * v))
            hi = helper._halo_from_next(v[0].contiguous())
            dz = D(v, hi)
            lo = helper._halo_from_prev(dz[0, -1].contiguous())
            w = A.back_project(A.project(v)) + D.adj(dz, lo)
            mu = gsum(v * w)
            v = w
        cnorm = math.sqrt(mu)
        # reference formula (_primaldual.py:286-288); note that factor > 1 loosens tau*sigma*||C||^2 < 1,
        # pass factor < 1 for a strict bound
        tau = math.sqrt(factor / ratio) / cnorm
        return tau, ratio * tau
