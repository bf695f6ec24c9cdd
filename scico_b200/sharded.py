"""Multi-GPU partitioning of the projector pair: one process per GPU over ``torch.distributed``.

Two partitionings:

* :class:`SlabShardedXRayTransform3D` -- z-slabs.  For separable geometry (detector rows depend on
  voxel axis 0 only: any rotation about axis 0) rank ``r`` holds volume slices ``[z0, z1)`` and the
  detector rows they project onto; the local plan gets ``slice_offset = z0``, the reference's own
  hook (``scico/linop/xray/_xray3d.py:143,195,208-212``).  No data-path collective, except a halo
  exchange of the detector rows two neighbouring slabs share when rows straddle a slab edge.
* :class:`ViewShardedXRayTransform3D` / :class:`ViewShardedXRayTransform2D` -- contiguous view
  blocks.  The forward needs the whole volume (``all_gather`` of the slab-sharded iterate); the
  adjoint's partial volumes are combined into slabs with a sum-reduction per destination slab,
  pipelined so that the back projection of slab ``j+1`` overlaps the reduction of slab ``j``.

The local operator is built by ``op_factory`` (default: the native CUDA operators of
:mod:`scico_b200.xray`).  Tests inject a CPU stand-in through it to exercise the partitioning and
the collectives under ``gloo``; the product never does.
"""

from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np

from . import geometry

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


class Partition:
    """What the reference's ``input_device`` / ``output_device`` arguments carry when they are a sharding rather
    than a device (``scico/linop/xray/_xray2d.py:60-61``, ``_xray3d.py:61-62,132,178``: the arrays are
    ``jax.device_put`` with it), for one process per GPU: ``XRayTransform3D(..., input_device=Partition("slabs"))``
    returns the z-slab operator of this rank, ``Partition("views")`` the view-block operator (volume in z-slabs,
    sinogram in view blocks), ``Partition("auto")`` z-slabs when the geometry is axis-0 separable and view blocks
    otherwise.  ``XRayTransform2D(..., output_device=Partition("views"))``: view blocks, image replicated.

    group / rank / world_size: the ``torch.distributed`` group (default: the world); exchange / rendezvous: how
    the view-block back projection meets its owners (``"nccl"``, ``"peer"``, ``"peer_add"``)."""

    def __init__(self, kind: str = "auto", group=None, rank: Optional[int] = None, world_size: Optional[int] = None,
                 exchange: str = "nccl", rendezvous: str = "flags"):
        if kind not in ("auto", "slabs", "views"):
            raise ValueError("Partition kind must be 'auto', 'slabs' or 'views'")
        self.kind, self.group, self.rank, self.world_size = kind, group, rank, world_size
        self.exchange, self.rendezvous = exchange, rendezvous

    def __repr__(self):
        return f"Partition({self.kind!r}, exchange={self.exchange!r})"


def partitioned_3d(part: "Partition", input_shape, matrices, det_shape):
    """The rank-local operator of ``XRayTransform3D`` under ``part`` (see :class:`Partition`)."""
    kind = part.kind
    if kind == "auto":
        kind = "slabs" if geometry.is_axis0_separable(np.asarray(matrices, dtype=np.float32)) else "views"
    if kind == "slabs":
        return SlabShardedXRayTransform3D(input_shape, matrices, det_shape, group=part.group, rank=part.rank,
                                          world_size=part.world_size)
    return ViewShardedXRayTransform3D(input_shape, matrices, det_shape, group=part.group, rank=part.rank,
                                      world_size=part.world_size, exchange=part.exchange, rendezvous=part.rendezvous)


def partitioned_2d(part: "Partition", input_shape, angles, **kw):
    """The rank-local operator of ``XRayTransform2D`` under ``part``: view blocks (the only 2D partition)."""
    if part.kind == "slabs":
        raise ValueError("a 2D projector has no z-slab partition: use Partition('views')")
    return ViewShardedXRayTransform2D(input_shape, angles, group=part.group, rank=part.rank, world_size=part.world_size,
                                      exchange=part.exchange, rendezvous=part.rendezvous, **kw)


def block_bounds(n: int, parts: int, index: int) -> tuple[int, int]:
    """Contiguous block ``index`` of ``n`` items split into ``parts`` nearly equal blocks."""
    if not 0 <= index < parts:
        raise ValueError(f"block index {index} outside [0, {parts})")
    return (n * index) // parts, (n * (index + 1)) // parts


def _world(group) -> tuple[int, int]:
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _native_3d(input_shape, matrices, det_shape, **kw):
    from .xray import XRayTransform3D

    return XRayTransform3D(input_shape, matrices, det_shape, **kw)


def _native_2d(input_shape, angles, **kw):
    from .xray import XRayTransform2D

    return XRayTransform2D(input_shape, angles, **kw)


class SlabShardedXRayTransform3D:
    """z-slab partition of ``XRayTransform3D`` for axis-0-separable geometry.

    Rank ``r`` owns volume slices ``slab = [z0, z1)`` and works on detector rows
    ``rows = [r0, r1)``.  ``project`` maps the local slab ``(z1-z0, N1, N2)`` to the local sinogram
    block ``(V, r1-r0, D1)``; ``back_project`` is its exact adjoint.  When neighbouring row ranges
    overlap (a voxel slice at a slab edge spreads over two detector rows) the shared rows are
    summed across the two ranks after ``project`` so that both hold the complete row, and
    ``owned_rows`` tells which of them this rank counts as its own (for inner products, norms and
    gathering): every global row is owned by exactly one rank.
    """

    def __init__(self, input_shape, matrices, det_shape, group=None, op_factory: Optional[Callable] = None,
                 rank: Optional[int] = None, world_size: Optional[int] = None):
        self.input_shape = tuple(int(s) for s in input_shape)
        self.det_shape = tuple(int(s) for s in det_shape)
        self.matrices = np.asarray(matrices, dtype=np.float32)
        if not geometry.is_axis0_separable(self.matrices):
            raise ValueError("z-slab sharding needs axis-0-separable matrices (use ViewShardedXRayTransform3D)")
        self.group = group
        r, w = _world(group)
        self.rank = r if rank is None else rank
        self.world_size = w if world_size is None else world_size
        n0 = self.input_shape[0]
        self.slabs = [block_bounds(n0, self.world_size, i) for i in range(self.world_size)]
        self.row_ranges = [geometry.slab_row_range(self.matrices, z0, z1, self.det_shape[0]) if z1 > z0 else (0, 0)
                           for z0, z1 in self.slabs]
        self.slab = self.slabs[self.rank]
        self.rows = self.row_ranges[self.rank]
        # ownership: a row held by several ranks (any number of them: with |M00| < 1/2 a detector row collects
        # slices of more than two slabs, and with M00 < 0 the row ranges run backwards) belongs to the
        # LOWEST rank that holds it; every global row a slab touches is then owned exactly once.
        self.owned_rows = self._ownership(self.rank)
        z0, z1 = self.slab
        self.local_input_shape = (z1 - z0,) + self.input_shape[1:]
        self.local_output_shape = (len(self.matrices), self.rows[1] - self.rows[0], self.det_shape[1])
        factory = op_factory or _native_3d
        self.local = None
        if z1 > z0 and self.rows[1] > self.rows[0]:
            self.local = factory(self.local_input_shape, self.matrices, self.local_output_shape[1:],
                                 slice_offset=z0, det_row_offset=self.rows[0], det_rows_total=self.det_shape[0])

    def _ownership(self, rank: int) -> tuple[int, int]:
        """Rows of ``row_ranges[rank]`` that no lower rank holds, as one interval ``(lo, hi)``."""
        lo, hi = self.row_ranges[rank]
        mine = np.ones(max(hi - lo, 0), dtype=bool)
        for q in range(rank):
            a, b = self.row_ranges[q]
            a, b = max(a, lo), min(b, hi)
            if b > a:
                mine[a - lo: b - lo] = False
        idx = np.flatnonzero(mine)
        if idx.size == 0:
            return (hi, hi)
        if idx[-1] - idx[0] + 1 != idx.size:
            # cannot happen for slabs of a monotone row map; refuse rather than drop rows from the sums
            raise ValueError(f"z-slab sharding: rank {rank}'s own detector rows are not one interval "
                             f"(row ranges {self.row_ranges}); use ViewShardedXRayTransform3D")
        return (lo + int(idx[0]), lo + int(idx[-1]) + 1)

    # -- halo: rows shared with other ranks -------------------------------------------------------
    def _overlap(self, other: int) -> tuple[int, int]:
        a, b = self.rows, self.row_ranges[other]
        return max(a[0], b[0]), min(a[1], b[1])

    def halo_rows(self) -> list[tuple[int, int, int]]:
        """[(other rank, global row lo, global row hi)] for every rank whose detector rows overlap this
        rank's -- the two neighbours for the usual |M00| >= 1/2, more ranks when a detector row collects
        the slices of several slabs."""
        out = []
        for other in range(self.world_size):
            if other == self.rank or self.slabs[other][1] <= self.slabs[other][0]:
                continue
            lo, hi = self._overlap(other)
            if hi > lo:
                out.append((other, lo, hi))
        return out

    def _exchange_halo_sum(self, y):
        halos = self.halo_rows()
        if not halos or self.world_size == 1:
            return y
        # NCCL moves device memory directly; gloo (the CPU / one-GPU test rendezvous) cannot send device pointers
        # (writev: "Bad address"), so the rows are staged through host memory there
        staged = y.is_cuda and dist.get_backend(self.group) == "gloo"
        ops, bufs = [], []
        for other, lo, hi in halos:
            send = y[:, lo - self.rows[0]:hi - self.rows[0], :].contiguous()
            if staged:
                send = send.cpu()
            recv = torch.empty_like(send)
            peer = other if self.group is None else dist.get_global_rank(self.group, other)
            ops += [dist.P2POp(dist.isend, send, peer, self.group), dist.P2POp(dist.irecv, recv, peer, self.group)]
            bufs.append((lo, hi, recv))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for lo, hi, recv in bufs:
            y[:, lo - self.rows[0]:hi - self.rows[0], :] += recv.to(y.device) if staged else recv
        return y

    def project(self, x_local, out=None):
        """``out``: optional preallocated local sinogram block (solvers reuse one buffer per iteration)."""
        if tuple(x_local.shape) != self.local_input_shape:
            raise ValueError(f"local slab of shape {tuple(x_local.shape)} does not match {self.local_input_shape}")
        if self.local is None:
            return x_local.new_zeros(self.local_output_shape) if out is None else out.zero_()
        return self._exchange_halo_sum(self.local.project(x_local, out=out) if out is not None else self.local.project(x_local))

    def back_project(self, y_local, out=None):
        if tuple(y_local.shape) != self.local_output_shape:
            raise ValueError(f"local sinogram of shape {tuple(y_local.shape)} does not match {self.local_output_shape}")
        if self.local is None:
            return y_local.new_zeros(self.local_input_shape) if out is None else out.zero_()
        return self.local.back_project(y_local, out=out) if out is not None else self.local.back_project(y_local)

    __call__ = project
    adj = back_project


class PeerBlocks:
    """Row blocks of a view-sharded back projection, held in device memory that every GPU of the node can
    write into (``xct_peer_alloc`` + CUDA IPC: one process per GPU, the buffers of the other ranks are mapped
    through NVLink / NVSwitch peer access).

    ``exchange(launch, out)`` runs one fused back projection + exchange step:

    1. ``launch(ptrs, row_begin, store)`` enqueues this rank's ``xct_adjoint_scatter``: the kernel's epilogue
       sends every result row to the rank that owns it --
       ``mode="store"`` (default): plain posted stores into THIS rank's slot of the owner's staging area
       ``(world, rows, *inner)``; ``mode="add"``: ``RED.ADD.F32`` (system scope) into the owner's block;
    2. one stream-ordered rendezvous: when it completes on this rank's stream, every rank's kernel has
       finished.  ``rendezvous="flags"`` (default): each rank writes the call's epoch into its flag word on
       every peer (``xct_peer_signal``, release / system scope, ordered after its kernel) and waits until all of
       its own words carry it (``xct_peer_wait``) -- two one-warp kernels, no collective library call;
       ``rendezvous="collective"``: a one-element all-reduce of the process group instead;
    3. store: the owner sums its slots in rank order into ``out`` (``xct_sum_slots``: deterministic, unlike
       a sum of atomics); add: the block is copied to ``out`` and zeroed again for its next use.

    Two copies of every buffer alternate between calls, which is what makes ONE rendezvous per call enough:
    a peer can only write copy ``c`` again two calls later, i.e. after a rendezvous that this rank joined
    after it had consumed ``c``.  No partial volume is written, sent and summed by a separate collective."""

    COPIES = 2

    def __init__(self, slabs: Sequence[tuple[int, int]], inner_shape: Sequence[int], group=None,
                 rank: Optional[int] = None, world_size: Optional[int] = None, device: Optional[int] = None,
                 mode: str = "store", mem=None, rendezvous: str = "flags", timeout_s: float = 30.0):
        if mode not in ("store", "add"):
            raise ValueError("mode must be 'store' or 'add'")
        if rendezvous not in ("flags", "collective"):
            raise ValueError("rendezvous must be 'flags' or 'collective'")
        self.mode = mode
        self.rendezvous = rendezvous
        self.timeout_s = float(timeout_s)
        self.group = group
        r, w = _world(group)
        self.rank = r if rank is None else rank
        self.world_size = w if world_size is None else world_size
        # `mem`: where the buffers live.  The product always uses the native CUDA-IPC memory; the CPU tests
        # inject a shared-memory stand-in to run this protocol (slot addressing, alternation) under gloo.
        self.mem = mem if mem is not None else _NativePeerMemory(device)
        from ._lib import MAX_ROUTE_PARTS

        if len(slabs) != self.world_size or self.world_size > MAX_ROUTE_PARTS:
            raise ValueError(f"need one row block per rank and at most {MAX_ROUTE_PARTS} ranks")
        self.slabs = [tuple(sl) for sl in slabs]
        self.row_begin = [self.slabs[0][0]] + [b for _, b in self.slabs]
        self.inner_shape = tuple(int(n) for n in inner_shape)
        inner = int(np.prod(self.inner_shape))
        self._block_elems = [(b - a) * inner for a, b in self.slabs]
        self.local_shape = (self.slabs[self.rank][1] - self.slabs[self.rank][0],) + self.inner_shape
        self.nelems = self._block_elems[self.rank]
        slots = self.world_size if mode == "store" else 1
        alloc = max(4 * self.nelems * slots, 256)  # an empty block still needs a valid pointer
        self._own: list = []
        self._mapped: list = []
        handles = []
        for _ in range(self.COPIES):
            ptr, handle = self.mem.alloc(alloc)
            # add: blocks start at zero; store: the slot of a rank without views is never written and stays zero
            self.mem.zero(ptr, alloc)
            self._own.append(ptr)
            handles.append(handle)
        # flag words of the rendezvous: COPIES x world int32, epochs only grow (never reset)
        self._flags_own, fh = self.mem.alloc(max(4 * self.COPIES * self.world_size, 256))
        self.mem.zero(self._flags_own, max(4 * self.COPIES * self.world_size, 256))
        handles.append(fh)
        self.mem.sync()
        gathered = [None] * self.world_size
        if self.world_size > 1:
            dist.all_gather_object(gathered, handles, group=group)
        else:
            gathered[0] = handles
        self.ptrs: list[list] = []  # [copy][owner] -> where this rank writes the owner's rows
        for c in range(self.COPIES):
            row = []
            for k in range(self.world_size):
                if k == self.rank:
                    base = self._own[c]
                else:
                    base = self.mem.open(gathered[k][c])
                    self._mapped.append(base)
                if mode == "store":  # this rank's slot in owner k's staging area
                    base = self.mem.offset(base, 4 * self.rank * self._block_elems[k])
                row.append(base)
            self.ptrs.append(row)
        # [copy][k]: this rank's flag word in rank k's flag array
        self._flag_ptrs: list[list] = []
        flag_bases = []
        for k in range(self.world_size):
            if k == self.rank:
                flag_bases.append(self._flags_own)
            else:
                q = self.mem.open(gathered[k][self.COPIES])
                self._mapped.append(q)
                flag_bases.append(q)
        for c in range(self.COPIES):
            self._flag_ptrs.append([self.mem.offset(flag_bases[k], 4 * (c * self.world_size + self.rank))
                                    for k in range(self.world_size)])
        self._epoch = [0] * self.COPIES
        self._token = self.mem.token()
        self._turn = 0
        self._closed = False
        if self.world_size > 1:
            dist.barrier(group=group)  # every buffer is zeroed and mapped before anybody writes into it

    def exchange(self, launch: Callable[[list, list, bool], None], out):
        if tuple(out.shape) != self.local_shape or not out.is_contiguous() or out.dtype != torch.float32:
            raise ValueError(f"'out' must be a contiguous float32 tensor of shape {self.local_shape}")
        c = self._turn
        self._turn = (c + 1) % self.COPIES
        launch(self.ptrs[c], self.row_begin, self.mode == "store")
        if self.world_size > 1:
            if self.rendezvous == "flags":  # stream-ordered, no host synchronisation, no collective call
                self._epoch[c] += 1
                self.mem.signal(self._flag_ptrs[c], self._epoch[c])
                self.mem.wait_flags(self.mem.offset(self._flags_own, 4 * c * self.world_size), self.world_size,
                                    self._epoch[c], self.timeout_s)
            else:
                dist.all_reduce(self._token, group=self.group)  # stream-ordered: no host synchronisation
        if self.nelems == 0:
            return out
        if self.mode == "store":
            self.mem.sum_slots(out, self._own[c], self.world_size, self.nelems)
        else:
            self.mem.copy_out(out, self._own[c], 4 * self.nelems)
            self.mem.zero(self._own[c], 4 * self.nelems)
        return out

    def close(self, collective: bool = True):
        """Unmap the peers' buffers and free this rank's.  MUST be called by every rank (``collective=True``,
        the default: two rendezvous make sure no peer is still writing into the buffers that are freed).
        The non-collective form exists for ``__del__`` / interpreter shutdown only: it waits for this rank's
        own device work and frees, and is safe only once the peers have stopped using the exchange."""
        if getattr(self, "_closed", True):
            return
        self._closed = True
        both = collective and self.world_size > 1 and dist.is_initialized()
        try:
            self.mem.sync()  # this rank's kernels (stores / REDs into mapped peer memory, slot sums) are done
        except Exception:  # interpreter shutdown
            pass
        if both:
            dist.barrier(group=self.group)
        timed_out = False
        try:  # a rendezvous that gave up on a peer has produced incomplete sums: say so, never silently
            timed_out = bool(getattr(self.mem, "timed_out", lambda: False)())
        except Exception:
            pass
        for q in self._mapped:
            self.mem.close(q)
        if both:
            dist.barrier(group=self.group)
        for q in self._own:
            self.mem.free(q)
        self.mem.free(self._flags_own)
        self._mapped, self._own = [], []
        if timed_out and collective:
            raise RuntimeError(f"PeerBlocks: a flag rendezvous timed out after {self.timeout_s} s (a peer rank stopped "
                               "taking part in the exchange); results since then are incomplete")

    def __del__(self):
        try:
            self.close(collective=False)
        except Exception:  # interpreter shutdown
            pass


class _NativePeerMemory:
    """Device buffers of :class:`PeerBlocks` through the C ABI (``xct_peer_*``, ``xct_sum_slots``): pointers
    are integers (device addresses), handles are the 64 bytes of a ``cudaIpcMemHandle_t``."""

    def __init__(self, device: Optional[int] = None):
        from . import _lib

        if torch is None or not torch.cuda.is_available():
            raise RuntimeError("PeerBlocks needs CUDA devices (there is no CPU path)")
        self._lib, self._L = _lib, _lib.lib()
        self.device = torch.cuda.current_device() if device is None else int(device)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def alloc(self, nbytes: int):
        import ctypes

        ptr, h = ctypes.c_void_p(), self._lib.IpcHandle()
        self._lib.check(self._L.xct_peer_alloc(self.device, nbytes, ctypes.byref(ptr), ctypes.byref(h)))
        return ptr.value, bytes(h.bytes)

    def open(self, handle: bytes):
        import ctypes

        h, q = self._lib.IpcHandle(), ctypes.c_void_p()
        ctypes.memmove(h.bytes, handle, 64)
        self._lib.check(self._L.xct_peer_open(self.device, ctypes.byref(h), ctypes.byref(q)))
        return q.value

    @staticmethod
    def offset(ptr: int, nbytes: int) -> int:
        return ptr + nbytes

    def zero(self, ptr: int, nbytes: int):
        self._lib.check(self._L.xct_peer_zero(self.device, ptr, nbytes, self._stream()))

    def sum_slots(self, out, ptr: int, nslots: int, n: int):
        self._lib.check(self._L.xct_sum_slots(self.device, out.data_ptr(), ptr, nslots, n, n, self._stream()))

    def copy_out(self, out, ptr: int, nbytes: int):
        self._lib.check(self._L.xct_peer_copy_out(self.device, out.data_ptr(), ptr, nbytes, self._stream()))

    def signal(self, flag_ptrs, epoch: int):
        import ctypes

        arr = (ctypes.c_void_p * len(flag_ptrs))(*flag_ptrs)
        self._lib.check(self._L.xct_peer_signal(self.device, arr, len(flag_ptrs), int(epoch), self._stream()))

    def wait_flags(self, flags_ptr: int, n: int, epoch: int, timeout_s: float):
        if not hasattr(self, "_timed_out"):
            self._timed_out = torch.zeros(1, dtype=torch.int32, device=f"cuda:{self.device}")
        self._lib.check(self._L.xct_peer_wait(self.device, flags_ptr, n, int(epoch), float(timeout_s),
                                              self._timed_out.data_ptr(), self._stream()))

    def timed_out(self) -> bool:
        """True if a rendezvous gave up waiting for a peer (host synchronisation; call when diagnosing)."""
        return hasattr(self, "_timed_out") and bool(self._timed_out.item())

    def token(self):
        return torch.zeros(1, dtype=torch.float32, device=f"cuda:{self.device}")

    def sync(self):
        torch.cuda.synchronize(self.device)

    def close(self, ptr: int):
        self._L.xct_peer_close(self.device, ptr)

    def free(self, ptr: int):
        self._L.xct_peer_free(self.device, ptr)


class _ViewSharded:
    """Shared machinery of the view-block partitions (volume / image rows sharded on axis 0)."""

    def _setup_exchange(self, exchange, inner_shape, peer_mem=None, rendezvous="flags", peer_timeout_s=30.0):
        if exchange not in ("nccl", "peer", "peer_add"):
            raise ValueError("exchange must be 'nccl', 'peer' (stores into per-rank slots) or 'peer_add' (atomics)")
        self.exchange = exchange
        self.peer = None
        if exchange != "nccl" and self.world_size > 1:
            self.peer = PeerBlocks(self.slabs, inner_shape, group=self.group, rank=self.rank, world_size=self.world_size,
                                   mode="store" if exchange == "peer" else "add", mem=peer_mem, rendezvous=rendezvous,
                                   timeout_s=peer_timeout_s)

    def close(self):
        if getattr(self, "peer", None) is not None:
            self.peer.close()
            self.peer = None

    def _setup(self, n_views, axis0, group, rank, world_size):
        self.group = group
        r, w = _world(group)
        self.rank = r if rank is None else rank
        self.world_size = w if world_size is None else world_size
        self.view_blocks = [block_bounds(n_views, self.world_size, i) for i in range(self.world_size)]
        self.views = self.view_blocks[self.rank]
        self.slabs = [block_bounds(axis0, self.world_size, i) for i in range(self.world_size)]
        self.slab = self.slabs[self.rank]

    def _gather_volume(self, x_slab):
        """all_gather of the axis-0 slabs into the full volume (uneven slabs are padded)."""
        if self.world_size == 1:
            return x_slab
        pad = max(z1 - z0 for z0, z1 in self.slabs)
        even = all(z1 - z0 == pad for z0, z1 in self.slabs)
        key = (tuple(x_slab.shape), x_slab.device, x_slab.dtype)
        if getattr(self, "_gather_key", None) != key:  # one padded send buffer and one gather buffer per operator
            self._gather_key = key
            self._gather_pad = None if even else x_slab.new_zeros((pad,) + tuple(x_slab.shape[1:]))
            self._gather_out = x_slab.new_empty((self.world_size * pad,) + tuple(x_slab.shape[1:]))
        if even:
            buf = x_slab.contiguous()
        else:
            buf = self._gather_pad
            buf[: x_slab.shape[0]] = x_slab
        out = self._gather_out
        dist.all_gather_into_tensor(out, buf, group=self.group)
        if all(z1 - z0 == pad for z0, z1 in self.slabs):
            return out
        return torch.cat([out[i * pad: i * pad + (z1 - z0)] for i, (z0, z1) in enumerate(self.slabs)], dim=0)

    def _reduce_slabs(self, partial_of_slab: Callable[[int], "torch.Tensor"]):
        """Sum the partial back projections of every rank, slab by slab, into the slab's owner.

        ``partial_of_slab(j)`` computes this rank's partial result for slab ``j``; the reduction of
        slab ``j`` is issued asynchronously, so it overlaps the computation of slab ``j+1``."""
        mine, works, keep = None, [], []
        for j in range(self.world_size):
            part = partial_of_slab(j)
            if self.world_size == 1:
                return part
            dst = j if self.group is None else dist.get_global_rank(self.group, j)
            works.append(dist.reduce(part, dst=dst, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            keep.append(part)
            if j == self.rank:
                mine = part
        for w in works:
            w.wait()
        return mine


class ViewShardedXRayTransform3D(_ViewSharded):
    """View-block partition of ``XRayTransform3D`` (any matrices).

    ``project(x_slab)``: all-gathers the slab-sharded volume and projects it onto this rank's views
    ``(v1-v0, D0, D1)``.  ``back_project(y_views)``: back-projects the local views slab by slab
    (one plan per destination slab, ``slice_offset`` = slab start) and sum-reduces each slab into
    its owner; returns this rank's slab ``(z1-z0, N1, N2)``.  ``exchange="peer"``: ONE back projection
    kernel whose epilogue writes every slice into its owner's memory through NVLink peer access
    (:class:`PeerBlocks`, ``xct_adjoint_scatter``; ``"peer_add"``: atomics instead of per-rank slots) instead
    of the per-slab NCCL reductions."""

    def __init__(self, input_shape, matrices, det_shape, group=None, op_factory: Optional[Callable] = None,
                 rank: Optional[int] = None, world_size: Optional[int] = None, exchange: str = "nccl", peer_mem=None,
                 rendezvous: str = "flags", peer_timeout_s: float = 30.0):
        self.input_shape = tuple(int(s) for s in input_shape)
        self.det_shape = tuple(int(s) for s in det_shape)
        self.matrices = np.asarray(matrices, dtype=np.float32)
        self._setup(len(self.matrices), self.input_shape[0], group, rank, world_size)
        self._setup_exchange(exchange, self.input_shape[1:], peer_mem, rendezvous, peer_timeout_s)  # peer_mem: CPU tests only
        v0, v1 = self.views
        self.local_output_shape = (v1 - v0,) + self.det_shape
        z0, z1 = self.slab
        self.local_input_shape = (z1 - z0,) + self.input_shape[1:]
        factory = op_factory or _native_3d
        M = self.matrices[v0:v1]
        self.full = factory(self.input_shape, M, self.det_shape) if v1 > v0 else None
        self.per_slab = [
            factory((b - a,) + self.input_shape[1:], M, self.det_shape, slice_offset=a) if (v1 > v0 and b > a) else None
            for a, b in self.slabs
        ]

    def project(self, x_slab, out=None):
        if tuple(x_slab.shape) != self.local_input_shape:
            raise ValueError(f"local slab of shape {tuple(x_slab.shape)} does not match {self.local_input_shape}")
        x = self._gather_volume(x_slab)
        if self.full is None:
            return x_slab.new_zeros(self.local_output_shape) if out is None else out.zero_()
        return self.full.project(x, out=out) if out is not None else self.full.project(x)

    def back_project(self, y_views, out=None):
        if tuple(y_views.shape) != self.local_output_shape:
            raise ValueError(f"local views of shape {tuple(y_views.shape)} do not match {self.local_output_shape}")

        if self.peer is not None:  # fused: the kernel adds every slice into its owner's slab over NVLink
            if out is None:
                out = y_views.new_empty(self.local_input_shape)
            launch = (lambda ptrs, rb, st: self.full.back_project_scatter(y_views, ptrs, rb, st)) if self.full is not None \
                else (lambda ptrs, rb, st: None)
            return self.peer.exchange(launch, out)

        def part(j):
            a, b = self.slabs[j]
            if self.per_slab[j] is None:
                return y_views.new_zeros((b - a,) + self.input_shape[1:])
            return self.per_slab[j].back_project(y_views)

        res = self._reduce_slabs(part)
        if out is not None:
            out.copy_(res)
            return out
        return res

    __call__ = project
    adj = back_project


class ViewShardedXRayTransform2D(_ViewSharded):
    """View-block partition of ``XRayTransform2D`` (BASELINE.json configs[2]: 4096^2, 2048 views).

    The image is small next to the sinogram, so it is kept replicated: ``project(x)`` maps the
    full image to this rank's views ``(v1-v0, ny)``; ``back_project(y_views)`` returns either the
    row block this rank owns (``scatter=True``, reduce per row block) or the full image on every
    rank (``scatter=False``, all-reduce).  ``exchange="peer"``: the scatter case runs ONE kernel whose
    epilogue writes every image row into its owner's memory through NVLink peer access
    (:class:`PeerBlocks`, ``xct_adjoint_scatter``; ``"peer_add"``: atomics instead of per-rank slots) instead of
    back projection + NCCL reduce-scatter."""

    def __init__(self, input_shape, angles, group=None, op_factory: Optional[Callable] = None,
                 rank: Optional[int] = None, world_size: Optional[int] = None, exchange: str = "nccl", peer_mem=None,
                 rendezvous: str = "flags", peer_timeout_s: float = 30.0, **kw):
        self.input_shape = tuple(int(s) for s in input_shape)
        self.angles = np.asarray(angles, dtype=np.float64)
        self._setup(len(self.angles), self.input_shape[0], group, rank, world_size)
        self._setup_exchange(exchange, self.input_shape[1:], peer_mem, rendezvous, peer_timeout_s)  # peer_mem: CPU tests only
        v0, v1 = self.views
        if kw.get("det_count") is None:  # the default depends on the image only, not on the views
            kw["det_count"] = int(np.ceil(np.linalg.norm(self.input_shape)))
        self.ny = int(kw["det_count"])
        self.local_output_shape = (v1 - v0, self.ny)
        factory = op_factory or _native_2d
        self.local = factory(self.input_shape, self.angles[v0:v1], **kw) if v1 > v0 else None

    def project(self, x, out=None):
        if tuple(x.shape) != self.input_shape:
            raise ValueError(f"image of shape {tuple(x.shape)} does not match {self.input_shape}")
        if self.local is None:
            return x.new_zeros(self.local_output_shape) if out is None else out.zero_()
        return self.local.project(x, out=out) if out is not None else self.local.project(x)

    def back_project(self, y_views, scatter: bool = True, out=None):
        res = self._back_project(y_views, scatter, out)
        if out is not None and res is not out:
            out.copy_(res)
            return out
        return res

    def _back_project(self, y_views, scatter, out):
        if tuple(y_views.shape) != self.local_output_shape:
            raise ValueError(f"local views of shape {tuple(y_views.shape)} do not match {self.local_output_shape}")
        if self.peer is not None and scatter:  # fused: image rows are added into their owners over NVLink
            if out is None:
                out = y_views.new_empty((self.slab[1] - self.slab[0],) + self.input_shape[1:])
            launch = (lambda ptrs, rb, st: self.local.back_project_scatter(y_views, ptrs, rb, st)) if self.local is not None \
                else (lambda ptrs, rb, st: None)
            return self.peer.exchange(launch, out)
        full = self.local.back_project(y_views) if self.local is not None else y_views.new_zeros(self.input_shape)
        if self.world_size == 1:
            return full
        if not scatter:
            dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
            return full
        even = len({b - a for a, b in self.slabs}) == 1
        if even and full.is_cuda and dist.get_backend(self.group) == "nccl":
            # equal row blocks: the partial image IS the reduce-scatter input, one NCCL call, no copies
            out = full.new_empty((self.slab[1] - self.slab[0],) + tuple(full.shape[1:]))
            dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=self.group)
            return out
        return self._reduce_slabs(lambda j: full[self.slabs[j][0]: self.slabs[j][1]].contiguous())

    __call__ = project
    adj = back_project
