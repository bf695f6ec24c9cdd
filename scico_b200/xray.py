"""``XRayTransform2D`` / ``XRayTransform3D``: drop-in mirrors of the reference operators
(``scico/linop/xray/_xray2d.py:29-156``, ``scico/linop/xray/_xray3d.py:27-108``) whose
``project`` / ``back_project`` run the hand-written sm_100a kernels through the C ABI.

Arrays: a CUDA ``torch.Tensor`` is processed in place on its device and current stream (no host
synchronisation); a NumPy array (or CPU tensor) goes through the host-buffer entry points
(H2D copy, kernels, D2H copy).  There is no CPU compute path.
"""

from __future__ import annotations

import ctypes
import threading
from typing import Optional
from warnings import warn

import numpy as np

from . import _lib
from .geometry import matrices_from_euler_angles, max_projected_width, view_table_2d
from .linop import LinearOperator

try:  # torch provides device memory and streams; it is plumbing, not the compute path
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_scalar_equiv(v) -> bool:
    return np.isscalar(v) or (hasattr(v, "ndim") and v.ndim == 0)


def _partition_of(input_device, output_device):
    """The :class:`scico_b200.sharded.Partition` among the two device arguments (None: plain devices)."""
    from .sharded import Partition

    parts = [d for d in (input_device, output_device) if isinstance(d, Partition)]
    if not parts:
        return None
    if len(parts) == 2 and parts[0] is not parts[1] and (parts[0].kind, parts[0].group) != (parts[1].kind, parts[1].group):
        raise ValueError("input_device and output_device name different partitions")
    return parts[0]


def _device_index(dev) -> Optional[int]:
    """Ordinal of an ``input_device`` / ``output_device`` argument (None = follow the input)."""
    if dev is None:
        return None
    if isinstance(dev, int):
        return dev
    if torch is not None:
        d = torch.device(dev)
        if d.type != "cuda":
            raise ValueError(f"scico_b200 operators run on CUDA devices only, got {dev!r}")
        return d.index if d.index is not None else torch.cuda.current_device()
    raise ValueError(f"cannot interpret device {dev!r}")


class _Plans:
    """Per-device native plans of one operator (created lazily, destroyed with the operator)."""

    def __init__(self, make):
        self._make = make
        self._plans: dict[int, ctypes.c_void_p] = {}
        self._lock = threading.Lock()
        # xct_*_host keeps its staging buffers, streams and events inside the plan (not re-entrant on one
        # plan): one lock per device serialises concurrent host-array applications of the same operator
        self._host_locks: dict[int, threading.Lock] = {}

    def get(self, device: int) -> ctypes.c_void_p:
        with self._lock:
            pl = self._plans.get(device)
            if pl is None:
                pl = self._make(device)
                self._plans[device] = pl
            return pl

    def host_lock(self, device: int) -> threading.Lock:
        with self._lock:
            return self._host_locks.setdefault(device, threading.Lock())

    def info(self, device: int) -> dict:
        inf = _lib.PlanInfo()
        _lib.check(_lib.lib().xct_plan_get_info(self.get(device), ctypes.byref(inf)))
        d = {name: getattr(inf, name) for name, _ in inf._fields_}
        d["path_name"] = _lib.PATH_NAMES.get(inf.path, "?")
        return d

    def close(self):
        with self._lock:
            for pl in self._plans.values():
                try:
                    _lib.lib().xct_plan_destroy(pl)
                except Exception:  # interpreter shutdown
                    pass
            self._plans.clear()

    def __del__(self):
        self.close()


def _apply(plans: _Plans, x, out_shape, forward: bool, batch: int, default_device: Optional[int], out=None,
           wait: bool = True):
    """Run one operator application on ``x`` (see module docstring for the array rules).

    ``out`` (optional): preallocated float32 C-contiguous result of shape ``out_shape`` -- a CUDA
    tensor for CUDA input, a NumPy array (ideally page-locked, e.g. the ``.numpy()`` view of a
    ``torch.empty(..., pin_memory=True)``) for host input.  It is overwritten and returned.

    ``wait=False`` (host arrays with ``out=`` only): enqueue and return; ``x`` and ``out`` must be page-locked and
    stay untouched until the operator's ``host_wait()`` returns (``xct_*_host_async`` / ``xct_host_wait``)."""
    L = _lib.lib()
    fn_dev = L.xct_forward if forward else L.xct_adjoint
    if wait:
        fn_host = L.xct_forward_host if forward else L.xct_adjoint_host
    else:
        fn_host = L.xct_forward_host_async if forward else L.xct_adjoint_host_async
    if torch is not None and isinstance(x, torch.Tensor) and x.is_cuda:
        dev = x.device.index
        xin = x.detach()
        if xin.dtype != torch.float32:
            xin = xin.to(torch.float32)
        xin = xin.contiguous()
        if out is None:
            out = torch.empty(out_shape, dtype=torch.float32, device=x.device)
        elif not (isinstance(out, torch.Tensor) and out.is_cuda and out.device == x.device and out.dtype == torch.float32
                  and tuple(out.shape) == tuple(out_shape) and out.is_contiguous()):
            raise ValueError("'out' must be a contiguous float32 CUDA tensor of the result shape on the input's device")
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(fn_dev(plans.get(dev), xin.data_ptr(), out.data_ptr(), batch, stream))
        if default_device is not None and default_device != dev:
            out = out.to(f"cuda:{default_device}")
        return out
    was_torch = torch is not None and isinstance(x, torch.Tensor)
    xin = np.ascontiguousarray(x.numpy() if was_torch else np.asarray(x), dtype=np.float32)
    if not wait and (out is None or xin is not (x.numpy() if was_torch else x)):
        raise ValueError("wait=False needs a float32 C-contiguous host input used as is and an explicit out= buffer")
    if out is None:
        out = np.empty(out_shape, dtype=np.float32)
    elif not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == tuple(out_shape)
              and out.flags.c_contiguous and out.flags.writeable):
        raise ValueError("'out' must be a writeable C-contiguous float32 NumPy array of the result shape")
    if default_device is not None:
        dev = default_device
    else:  # host arrays run on the process's current CUDA device (one process per GPU under torchrun)
        dev = torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0
    with plans.host_lock(dev):
        _lib.check(fn_host(plans.get(dev), xin.ctypes.data, out.ctypes.data, batch))
    return torch.from_numpy(out) if was_torch else out


if torch is not None:

    class _ProjectorFn(torch.autograd.Function):
        """The pair as one differentiable linear map for ``torch.autograd``: the VJP of ``project`` is
        ``back_project`` and vice versa, which is how the reference wires an external projector
        (``jax.custom_vjp(self._proj)`` with ``bwd -> self._bproj``, ``scico/linop/xray/astra/_astra_3d.py:498-502``)
        and what ``test/linop/xray/astra/test_astra_2d.py:145-187`` checks (``grad ||A x||^2 == 2 A^T A x``, gradient
        through ``A.T``).  The backward pass is differentiable again (it is the other kernel)."""

        @staticmethod
        def forward(ctx, x, plans, out_shape, forward, batch, default_device):
            ctx.args = (plans, tuple(x.shape), forward, batch, x.device)
            return _apply(plans, x, out_shape, forward, batch, default_device)

        @staticmethod
        def backward(ctx, g):
            plans, in_shape, forward, batch, x_device = ctx.args
            gx = _apply_ad(plans, g.contiguous(), in_shape, not forward, batch, None)
            if gx.device != x_device:
                gx = gx.to(x_device)
            return gx, None, None, None, None, None


def _apply_ad(plans: _Plans, x, out_shape, forward: bool, batch: int, default_device: Optional[int], out=None,
              wait: bool = True):
    """:func:`_apply`, recorded on the autograd tape when ``x`` is a tensor that requires grad."""
    if torch is not None and isinstance(x, torch.Tensor) and x.requires_grad and torch.is_grad_enabled():
        if out is not None:  # writing into a caller's buffer cannot be recorded: refuse rather than detach silently
            raise ValueError("'out=' cannot be combined with an input that requires grad (use torch.no_grad() or drop out=)")
        return _ProjectorFn.apply(x, plans, tuple(out_shape), forward, batch, default_device)
    return _apply(plans, x, out_shape, forward, batch, default_device, out, wait)


def _host_wait(plans: _Plans, device: Optional[int]):
    """Block until every ``wait=False`` host-array application of this operator has delivered its result."""
    dev = device if device is not None else (torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0)
    with plans.host_lock(dev):
        _lib.check(_lib.lib().xct_host_wait(plans.get(dev)))


def _analyse(fn, geom) -> dict:
    info, cls = _lib.PlanInfo(), _lib.PlanClasses()
    _lib.check(fn(ctypes.byref(geom), ctypes.byref(info), ctypes.byref(cls)))
    d = {name: getattr(info, name) for name, _ in info._fields_}
    d["path_name"] = _lib.PATH_NAMES.get(info.path, "?")
    d["joint_views"] = list(cls.joint_views)
    d["two_bin_views"] = list(cls.two_bin_views)
    d["brick_views"] = list(cls.brick_views)
    for name in ("adj_jump_views", "rows_unit", "rows_consecutive", "fwd_cold", "fwd_tile", "adj_interleaved"):
        d[name] = getattr(cls, name)
    return d


def _apply_scatter(plans: _Plans, y, ptrs, row_begin, store: bool = False):
    """Back projection of the CUDA tensor ``y`` whose result rows go to the row blocks behind ``ptrs``
    (device pointers, local or peer-mapped; block ``k`` = rows ``[row_begin[k], row_begin[k+1])`` of axis 0):
    ``xct_adjoint_scatter``.  ``store=False``: the values are ADDED (system-scope RED); ``store=True``: they
    are stored, each element exactly once.  Asynchronous on the current stream."""
    if not (torch is not None and isinstance(y, torch.Tensor) and y.is_cuda):
        raise ValueError("back_project_scatter needs a CUDA tensor")
    if len(ptrs) + 1 != len(row_begin) or not 1 <= len(ptrs) <= _lib.MAX_ROUTE_PARTS:
        raise ValueError(f"need 1..{_lib.MAX_ROUTE_PARTS} row blocks and one more row boundary than blocks")
    route = _lib.OutRoute()
    route.nparts = len(ptrs)
    route.store = 1 if store else 0
    for k, b in enumerate(row_begin):
        route.row_begin[k] = int(b)
    for k, q in enumerate(ptrs):
        route.ptr[k] = int(q) if q else None
    dev = y.device.index
    yin = y.detach()
    if yin.dtype != torch.float32:
        yin = yin.to(torch.float32)
    yin = yin.contiguous()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().xct_adjoint_scatter(plans.get(dev), yin.data_ptr(), ctypes.byref(route), stream))


class XRayTransform2D(LinearOperator):
    r"""Parallel ray, single axis, 2D X-ray projector (``_xray2d.py:29-136``).

    Same constructor arguments, defaults, public attributes and warning behaviour as the
    reference.  `x0`, `dx` and `y0` are in units of the detector spacing ``dy = 1``.
    """

    def __new__(cls, input_shape=None, angles=None, x0=None, dx=None, y0=None, det_count=None,
                input_device=None, output_device=None, _flags: int = 0):
        # a sharding in place of a device (reference: jax.device_put(x, sharding), _xray2d.py:60-61,257,298):
        # the constructor hands back this rank's partitioned operator (scico_b200.sharded.Partition)
        part = _partition_of(input_device, output_device)
        if part is not None:
            from .sharded import partitioned_2d

            kw = {k: v for k, v in dict(x0=x0, dx=dx, y0=y0, det_count=det_count).items() if v is not None}
            return partitioned_2d(part, input_shape, angles, **kw)
        return super().__new__(cls)

    def __init__(self, input_shape, angles, x0=None, dx=None, y0=None, det_count=None,
                 input_device=None, output_device=None, _flags: int = 0):
        self.input_shape = tuple(input_shape)
        self.angles = angles
        self.nx = tuple(input_shape)
        if dx is None:
            dx = 2 * (np.sqrt(2) / 2,)
        if _is_scalar_equiv(dx):
            dx = 2 * (dx,)
        self.dx = dx

        max_width = max_projected_width(np.asarray(angles), (float(dx[0]), float(dx[1])))
        if max_width > 1:
            warn(f"A projected pixel has width {max_width} > 1.0, "
                 "which will reduce projector accuracy.")

        if x0 is None:
            x0 = -(np.array(self.nx) * np.asarray(self.dx, dtype=np.float64)) / 2
        self.x0 = x0
        if det_count is None:
            det_count = int(np.ceil(np.linalg.norm(input_shape)))
        self.det_count = det_count
        self.ny = det_count
        self.output_shape = (len(angles), det_count)
        if y0 is None:
            y0 = -self.ny / 2
        self.y0 = y0
        self.dy = 1.0
        self.fbp_filter = None
        self.fbp_mask = None
        self.input_device = input_device
        self.output_device = output_device

        self.view_table = view_table_2d(np.asarray(angles), self.x0, self.dx, self.y0)
        self._flags = _flags
        self._plans = _Plans(self._make_plan)

        super().__init__(input_shape=self.input_shape, input_dtype=np.float32,
                         output_shape=self.output_shape, output_dtype=np.float32,
                         eval_fn=self.project, adj_fn=self.back_project)

    def _geom(self, device: int):
        g = _lib.Geom2D()
        g.n0, g.n1 = int(self.nx[0]), int(self.nx[1])
        g.num_views = int(self.view_table.shape[0])
        g.det_count = int(self.ny)
        g.view_table = self.view_table.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        g.device = device
        g.flags = self._flags
        return g

    def _make_plan(self, device: int):
        g = self._geom(device)
        pl = ctypes.c_void_p()
        _lib.check(_lib.lib().xct2d_plan_create(ctypes.byref(pl), ctypes.byref(g)))
        return pl

    def analyse(self) -> dict:
        """The native plan's decisions for this geometry (kernel families, view classes), computed on the
        host without a CUDA device (``xct2d_plan_analyse``)."""
        return _analyse(_lib.lib().xct2d_plan_analyse, self._geom(0))

    def plan_info(self, device: int = 0) -> dict:
        return self._plans.info(device)

    def project(self, im, out=None):
        """X-ray projection, ``H @ im``; a leading batch axis is accepted (the reference's
        ``jax.vmap(A)`` use, ``scico/flax/examples/data_generation.py:153-186``)."""
        batch, lead = self._batch(im, self.nx)
        return _apply_ad(self._plans, im, lead + self.output_shape, True, batch,
                      _device_index(self.output_device), out)

    def back_project(self, y, out=None):
        """X-ray back projection, ``H.T @ y`` (exact adjoint of :meth:`project`)."""
        batch, lead = self._batch(y, self.output_shape)
        return _apply_ad(self._plans, y, lead + self.nx, False, batch, _device_index(self.input_device), out)

    def back_project_scatter(self, y, ptrs, row_begin, store: bool = False) -> None:
        """Back projection fused with the view-block exchange: image row ``i`` is added into (or, with
        ``store``, written to) the row block that owns it (``ptrs[k]``: device pointer of rows
        ``[row_begin[k], row_begin[k+1])``, local or a peer GPU's, see :class:`scico_b200.sharded.PeerBlocks`).
        One image, CUDA tensors only."""
        if tuple(y.shape) != self.output_shape:
            raise ValueError(f"array of shape {tuple(y.shape)} does not match {self.output_shape}")
        _apply_scatter(self._plans, y, ptrs, row_begin, store)

    @staticmethod
    def _batch(a, core):
        shp = tuple(a.shape)
        if shp == tuple(core):
            return 1, ()
        if len(shp) == len(core) + 1 and shp[1:] == tuple(core):
            return shp[0], (shp[0],)
        raise ValueError(f"array of shape {shp} does not match operator shape {tuple(core)}")

    def fbp(self, y):
        r"""Filtered back projection (``_xray2d.py:158-197``): ramp filter of Kak & Slaney eq. 61
        applied in the frequency domain with padding to ``2N-1``, then back projection, masked to
        the pixels every view sees and scaled by :math:`\pi\,dx_0 dx_1 / V`."""
        N = y.shape[1]
        is_t = torch is not None and isinstance(y, torch.Tensor)
        if self.fbp_filter is None:
            n = np.arange(N) - (N - 1) // 2
            h = np.where(n == 0, 0.25, np.where(n % 2, -1.0 / (n.astype(np.float64) ** 2 * np.pi**2 + (n == 0)), 0.0))
            self.fbp_filter = h.astype(np.float32).reshape(1, -1)
        if self.fbp_mask is None:
            ones = torch.ones_like(y) if is_t else np.ones_like(y)
            self.fbp_mask = self.back_project(ones) >= (self.output_shape[0] * (1.0 - 1e-5))
        L = 2 * N - 1
        lo, hi = (N - 1) // 2, -(N - 1) // 2
        if is_t:
            h = torch.as_tensor(self.fbp_filter, device=y.device)
            hy = torch.fft.ifft(torch.fft.fft(h, n=L, dim=1) * torch.fft.fft(y, n=L, dim=1), n=L, dim=1)
            hy = hy[:, lo:hi].real.to(torch.float32).contiguous()
            mask = self.fbp_mask.to(y.device) if isinstance(self.fbp_mask, torch.Tensor) else torch.as_tensor(self.fbp_mask, device=y.device)
        else:
            hy = np.fft.ifft(np.fft.fft(self.fbp_filter, n=L, axis=1) * np.fft.fft(y, n=L, axis=1), n=L, axis=1)
            hy = np.ascontiguousarray(hy[:, lo:hi].real, dtype=np.float32)
            mask = self.fbp_mask if isinstance(self.fbp_mask, np.ndarray) else self.fbp_mask.cpu().numpy()
        scale = np.float32(np.pi * self.dx[0] * self.dx[1] / y.shape[0])
        return scale * mask * self.back_project(hy)


class XRayTransform3D(LinearOperator):
    r"""General-purpose, 3D, parallel ray X-ray projector (``_xray3d.py:27-108``).

    One (2, 4) homogeneous matrix per view maps the voxel centre ``(i+.5, j+.5, k+.5)`` to
    detector coordinates; the detector pixel ``(r, c)`` covers ``[r, r+1) x [c, c+1)``.
    ``slice_offset`` / ``det_row_offset`` / ``det_rows_total`` expose the reference's unused
    z-slab hook (``_xray3d.py:143,195,208-212``) for multi-GPU sharding.
    """

    def __new__(cls, input_shape=None, matrices=None, det_shape=None, batch_size: int = 8,
                input_dtype=np.float32, input_device=None, output_device=None, **kw):
        # a sharding in place of a device (reference: _xray3d.py:61-62,132,178): this rank's partitioned operator
        part = _partition_of(input_device, output_device)
        if part is not None:
            from .sharded import partitioned_3d

            if kw:
                raise ValueError(f"{sorted(kw)} cannot be combined with a Partition (the partition sets the offsets)")
            return partitioned_3d(part, input_shape, matrices, det_shape)
        return super().__new__(cls)

    def __init__(self, input_shape, matrices, det_shape, batch_size: int = 8,
                 input_dtype=np.float32, input_device=None, output_device=None, *,
                 slice_offset: int = 0, det_row_offset: int = 0, det_rows_total: int = 0,
                 _flags: int = 0):
        self.input_shape = tuple(input_shape)
        self.matrices = np.ascontiguousarray(np.asarray(matrices), dtype=np.float32)
        if self.matrices.ndim != 3 or self.matrices.shape[1:] != (2, 4):
            raise ValueError(f"matrices must have shape (num_views, 2, 4), got {self.matrices.shape}")
        if np.dtype(input_dtype) != np.float32:
            raise ValueError("scico_b200 XRayTransform3D computes in float32 only")
        self.det_shape = tuple(det_shape)
        self.batch_size = batch_size  # accepted for API compatibility; views are not batched here
        self.output_shape = (len(self.matrices), *self.det_shape)
        self.input_device = input_device
        self.output_device = output_device
        self.slice_offset = int(slice_offset)
        self.det_row_offset = int(det_row_offset)
        self.det_rows_total = int(det_rows_total)
        self._flags = _flags
        self._plans = _Plans(self._make_plan)
        super().__init__(input_shape=self.input_shape, output_shape=self.output_shape,
                         eval_fn=self.project, adj_fn=self.back_project,
                         input_dtype=input_dtype, output_dtype=input_dtype)

    def _geom(self, device: int):
        g = _lib.Geom3D()
        g.n0, g.n1, g.n2 = (int(s) for s in self.input_shape)
        g.d0, g.d1 = (int(s) for s in self.det_shape)
        g.num_views = int(self.matrices.shape[0])
        g.matrices = self.matrices.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        g.slice_offset = self.slice_offset
        g.det_row_offset = self.det_row_offset
        g.det_rows_total = self.det_rows_total
        g.device = device
        g.flags = self._flags
        return g

    def _make_plan(self, device: int):
        g = self._geom(device)
        pl = ctypes.c_void_p()
        _lib.check(_lib.lib().xct3d_plan_create(ctypes.byref(pl), ctypes.byref(g)))
        return pl

    def analyse(self) -> dict:
        """The native plan's decisions for this geometry (kernel families, view classes, TMA / joint
        eligibility), computed on the host without a CUDA device (``xct3d_plan_analyse``)."""
        return _analyse(_lib.lib().xct3d_plan_analyse, self._geom(0))

    def plan_info(self, device: int = 0) -> dict:
        return self._plans.info(device)

    def project(self, im, out=None, wait: bool = True):
        """Compute X-ray projection (``out``: optional preallocated result; ``wait=False``: host arrays only,
        enqueue and return, see :func:`_apply` and :meth:`host_wait`)."""
        if tuple(im.shape) != self.input_shape:
            raise ValueError(f"array of shape {tuple(im.shape)} does not match {self.input_shape}")
        return _apply_ad(self._plans, im, self.output_shape, True, 1, _device_index(self.output_device), out, wait)

    def back_project(self, proj, out=None, wait: bool = True):
        """Compute X-ray back projection (exact adjoint of :meth:`project`)."""
        if tuple(proj.shape) != self.output_shape:
            raise ValueError(f"array of shape {tuple(proj.shape)} does not match {self.output_shape}")
        return _apply_ad(self._plans, proj, self.input_shape, False, 1, _device_index(self.input_device), out, wait)

    def host_wait(self, device: Optional[int] = None):
        """Wait for the ``wait=False`` applications of this operator on host arrays (``xct_host_wait``)."""
        _host_wait(self._plans, device)

    def back_project_scatter(self, proj, ptrs, row_begin, store: bool = False) -> None:
        """Back projection fused with the view-block exchange: slice ``i`` of the result is added into (or,
        with ``store``, written to) the slab that owns it (``ptrs[k]``: device pointer of slices
        ``[row_begin[k], row_begin[k+1])``, local or a peer GPU's, see :class:`scico_b200.sharded.PeerBlocks`).
        CUDA tensors only."""
        if tuple(proj.shape) != self.output_shape:
            raise ValueError(f"array of shape {tuple(proj.shape)} does not match {self.output_shape}")
        _apply_scatter(self._plans, proj, ptrs, row_begin, store)

    matrices_from_euler_angles = staticmethod(matrices_from_euler_angles)


def debug_weights_3d(op: XRayTransform3D, view: int, device: int = 0):
    """(ul (2,...) int32, w (4,...) f32) as computed by the plan's kernels (test hook)."""
    n = op.input_shape
    ul = torch.empty((2, *n), dtype=torch.int32, device=f"cuda:{device}")
    w = torch.empty((4, *n), dtype=torch.float32, device=f"cuda:{device}")
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        _lib.check(_lib.lib().xct3d_debug_weights(op._plans.get(device), view, ul.data_ptr(), w.data_ptr(), st))
    return ul.cpu().numpy(), w.cpu().numpy()


def debug_weights_2d(op: XRayTransform2D, view: int, device: int = 0):
    """(inds (N0,N1) int32, w (N0,N1) f32) as computed by the kernels (test hook)."""
    inds = torch.empty(op.nx, dtype=torch.int32, device=f"cuda:{device}")
    w = torch.empty(op.nx, dtype=torch.float32, device=f"cuda:{device}")
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        _lib.check(_lib.lib().xct2d_debug_weights(op._plans.get(device), view, inds.data_ptr(), w.data_ptr(), st))
    return inds.cpu().numpy(), w.cpu().numpy()
