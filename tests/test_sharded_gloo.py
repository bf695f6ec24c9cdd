"""Multi-process (world_size 2, gloo, CPU) tests of the partitioning + collective plumbing in
scico_b200/sharded.py.  The local operator is a stand-in built on the oracle's C port (injected
through `op_factory`; test infrastructure only), so what is tested here is the host logic: slab /
row / view bounds, the halo exchange, the all-gather and the pipelined slab reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scico_b200 import geometry, sharded
from oracle import xray_c as C
from oracle import xray_np as O


class OracleOp3D:
    """CPU stand-in with the constructor / methods of scico_b200.XRayTransform3D."""

    def __init__(self, input_shape, matrices, det_shape, slice_offset=0, det_row_offset=0, det_rows_total=0):
        self.N, self.M, self.D = tuple(input_shape), np.asarray(matrices, np.float32), tuple(det_shape)
        self.so, self.ro = slice_offset, det_row_offset
        self.rt = det_rows_total or self.D[0]

    def project(self, x):
        y = C.project_3d(x.numpy(), self.M, (self.rt, self.D[1]), slice_offset=self.so)
        return torch.from_numpy(np.ascontiguousarray(y[:, self.ro:self.ro + self.D[0]]))

    def back_project(self, y):
        full = np.zeros((len(self.M), self.rt, self.D[1]), np.float32)
        full[:, self.ro:self.ro + self.D[0]] = y.numpy()
        return torch.from_numpy(C.back_project_3d(full, self.M, self.N, slice_offset=self.so))


class OracleOp2D:
    def __init__(self, input_shape, angles, det_count=None, **kw):
        self.nx, self.ny = tuple(input_shape), int(det_count)
        dx = 2 * (np.sqrt(2) / 2,)
        x0 = -(np.array(self.nx) * np.asarray(dx)) / 2
        self.T = O.view_table_2d(angles, x0, dx, -self.ny / 2)

    def project(self, x):
        return torch.from_numpy(C.project_2d(x.numpy(), self.T, self.ny))

    def back_project(self, y):
        return torch.from_numpy(C.back_project_2d(y.numpy(), self.T, self.nx))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(rank, world, port, fn, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def _spawn(fn, *args, world=2):
    mp.spawn(_run, args=(world, _free_port(), fn, args), nprocs=world, join=True)


def _gather_rows(op, y_local, V, D):
    """Assemble the global sinogram from the rows every rank owns."""
    full = torch.zeros((V,) + D)
    lo, hi = op.owned_rows
    full[:, lo:hi] = y_local[:, lo - op.rows[0]: hi - op.rows[0]]
    dist.all_reduce(full)
    return full


# ---------------------------------------------------------------------------------------------
def _slab_case(rank, world, N, D, V, spacing, flip=False):
    M = O.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None], voxel_spacing=spacing)
    if flip:  # detector rows run against the slice index: M00 < 0
        M = np.array(M, copy=True)
        M[:, 0, 0], M[:, 0, 3] = -M[:, 0, 0], D[0] - M[:, 0, 3]
    rng = np.random.default_rng(5)
    x = rng.standard_normal(N).astype(np.float32)
    y = rng.standard_normal((V,) + D).astype(np.float32)
    op = sharded.SlabShardedXRayTransform3D(N, M, D, op_factory=OracleOp3D)
    assert op.rank == rank and op.world_size == world
    z0, z1 = op.slab
    got = op.project(torch.from_numpy(x[z0:z1].copy()))
    assert tuple(got.shape) == op.local_output_shape
    full = _gather_rows(op, got, V, D).numpy()
    want = C.project_3d(x, M.astype(np.float32), D)
    assert O.rel_l2(full, want) <= 1e-6
    r0, r1 = op.rows
    back = op.back_project(torch.from_numpy(np.ascontiguousarray(y[:, r0:r1]))).numpy()
    want_b = C.back_project_3d(y, M.astype(np.float32), N)[z0:z1]
    assert O.rel_l2(back, want_b) <= 1e-6
    # adjoint identity with ownership-aware inner products
    lo, hi = op.owned_rows
    a = torch.tensor(float(np.sum(got.numpy()[:, lo - r0:hi - r0].astype(np.float64) * y[:, lo:hi])))
    b = torch.tensor(float(np.sum(back.astype(np.float64) * x[z0:z1])))
    dist.all_reduce(a)
    dist.all_reduce(b)
    assert abs(a.item() - b.item()) / max(abs(a.item()), abs(b.item())) < 1e-5


def test_slab_sharded_aligned_rows():
    _spawn(_slab_case, (12, 10, 9), (12, 14), 5, None)


def test_slab_sharded_with_row_halo():
    """voxel spacing 0.8 along axis 0: slices at the slab edge spread over a shared detector row."""
    _spawn(_slab_case, (11, 10, 9), (12, 14), 4, [0.8, 1.0, 1.0])


def test_slab_sharded_three_ranks_uneven():
    _spawn(_slab_case, (10, 8, 9), (10, 12), 3, None, world=3)


def test_slab_sharded_rows_shared_by_more_than_two_ranks():
    """|M00| = 0.4 with one-slice slabs: a detector row collects the slices of three or four ranks, so the
    halo exchange and the row ownership must look beyond the two neighbours."""
    _spawn(_slab_case, (4, 6, 5), (6, 8), 3, [0.4, 1.0, 1.0], world=4)


def test_slab_sharded_reversed_row_order():
    """M00 < 0: row ranges decrease with the rank; every touched row must still be owned exactly once."""
    _spawn(_slab_case, (11, 10, 9), (12, 14), 4, [0.8, 1.0, 1.0], True)


def _view_case_3d(rank, world, N, D, V, seq):
    ang = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, 0.4)], 1) if seq == "XY" else \
        np.linspace(0, np.pi, V, endpoint=False)[:, None]
    M = O.matrices_from_euler_angles(N, D, seq, ang)
    rng = np.random.default_rng(6)
    x = rng.standard_normal(N).astype(np.float32)
    y = rng.standard_normal((V,) + D).astype(np.float32)
    op = sharded.ViewShardedXRayTransform3D(N, M, D, op_factory=OracleOp3D)
    z0, z1 = op.slab
    v0, v1 = op.views
    got = op.project(torch.from_numpy(x[z0:z1].copy())).numpy()
    want = C.project_3d(x, M.astype(np.float32), D)[v0:v1]
    assert O.rel_l2(got, want) <= 1e-6
    back = op.back_project(torch.from_numpy(np.ascontiguousarray(y[v0:v1]))).numpy()
    want_b = C.back_project_3d(y, M.astype(np.float32), N)[z0:z1]
    assert back.shape == want_b.shape and O.rel_l2(back, want_b) <= 1e-6


def test_view_sharded_3d_tilted_geometry():
    _spawn(_view_case_3d, (9, 8, 7), (12, 11), 5, "XY")


def test_view_sharded_3d_uneven_three_ranks():
    _spawn(_view_case_3d, (10, 6, 7), (10, 10), 7, "X", world=3)


def _view_case_2d(rank, world, nx, V):
    angles = np.linspace(0, np.pi, V, endpoint=False)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(nx).astype(np.float32)
    op = sharded.ViewShardedXRayTransform2D(nx, angles, op_factory=OracleOp2D)
    y = rng.standard_normal((V, op.ny)).astype(np.float32)
    ref = OracleOp2D(nx, angles, det_count=op.ny)
    v0, v1 = op.views
    got = op.project(torch.from_numpy(x)).numpy()
    assert O.rel_l2(got, ref.project(torch.from_numpy(x)).numpy()[v0:v1]) <= 1e-6
    want_b = ref.back_project(torch.from_numpy(y)).numpy()
    z0, z1 = op.slab
    blk = op.back_project(torch.from_numpy(np.ascontiguousarray(y[v0:v1]))).numpy()
    assert O.rel_l2(blk, want_b[z0:z1]) <= 1e-5
    rep = op.back_project(torch.from_numpy(np.ascontiguousarray(y[v0:v1])), scatter=False).numpy()
    assert O.rel_l2(rep, want_b) <= 1e-5


def test_view_sharded_2d():
    _spawn(_view_case_2d, (24, 20), 9)


# ---------------------------------------------------------------------------------------------
# Fused view-block exchange (sharded.PeerBlocks): the protocol -- handle exchange, slot addressing with
# uneven row blocks, a rank without views, the two alternating copies, the slot sum in rank order -- run
# under gloo with shared memory standing in for the CUDA-IPC buffers and the oracle for the routed kernel.
class ShmPeerMemory:
    """multiprocessing.shared_memory stand-in for sharded._NativePeerMemory: pointer = (segment name, byte
    offset), handle = the segment name."""

    def __init__(self):
        self._segs, self._mine = {}, {}

    def _seg(self, name):
        from multiprocessing import shared_memory

        if name not in self._segs:
            self._segs[name] = shared_memory.SharedMemory(name=name)
        return self._segs[name]

    def view(self, ptr, n):
        return np.ndarray((n,), np.float32, buffer=self._seg(ptr[0]).buf, offset=ptr[1])

    def alloc(self, nbytes):
        from multiprocessing import shared_memory

        shm = shared_memory.SharedMemory(create=True, size=nbytes)
        self._segs[shm.name] = self._mine[shm.name] = shm
        return (shm.name, 0), shm.name.encode()

    def open(self, handle):
        self._seg(handle.decode())
        return (handle.decode(), 0)

    @staticmethod
    def offset(ptr, nbytes):
        return (ptr[0], ptr[1] + nbytes)

    def zero(self, ptr, nbytes):
        self.view(ptr, nbytes // 4)[:] = 0

    def sum_slots(self, out, ptr, nslots, n):
        slots = self.view(ptr, nslots * n).reshape(nslots, n)
        acc = slots[0].copy()
        for s in range(1, nslots):
            acc = acc + slots[s]
        out.numpy().reshape(-1)[:] = acc

    def copy_out(self, out, ptr, nbytes):
        out.numpy().reshape(-1)[:] = self.view(ptr, nbytes // 4)

    def signal(self, flag_ptrs, epoch):
        for q in flag_ptrs:
            np.ndarray((1,), np.int32, buffer=self._seg(q[0]).buf, offset=q[1])[0] = epoch

    def wait_flags(self, flags_ptr, n, epoch, timeout_s):
        import time

        words = np.ndarray((n,), np.int32, buffer=self._seg(flags_ptr[0]).buf, offset=flags_ptr[1])
        t0 = time.time()
        while not np.all(words >= epoch):
            assert time.time() - t0 < timeout_s, "rendezvous timed out"
            time.sleep(1e-4)

    @staticmethod
    def token():
        return torch.zeros(1)

    def sync(self):
        pass

    def close(self, ptr):
        if ptr[0] not in self._mine:
            self._segs.pop(ptr[0]).close()

    def free(self, ptr):
        shm = self._mine.pop(ptr[0])
        self._segs.pop(ptr[0], None)
        shm.close()
        shm.unlink()


_MEM = None  # the worker's ShmPeerMemory (the stand-in operators below write through it)


class _ScatterMixin:
    """back_project_scatter of the native operators, on the shared-memory stand-in."""

    def back_project_scatter(self, y, ptrs, row_begin, store=False):
        full = self.back_project(y).numpy()
        inner = int(np.prod(full.shape[1:]))

        def write():
            for k, (a, b) in enumerate(zip(row_begin[:-1], row_begin[1:])):
                if b > a:
                    dst = _MEM.view(ptrs[k], (b - a) * inner)
                    if store:
                        dst[:] = full[a:b].reshape(-1)
                    else:
                        dst += full[a:b].reshape(-1)

        if store:
            write()
        else:  # the GPUs add atomically; the stand-in takes turns
            for turn in range(dist.get_world_size()):
                if turn == dist.get_rank():
                    write()
                dist.barrier()


class ScatterOp2D(_ScatterMixin, OracleOp2D):
    pass


class ScatterOp3D(_ScatterMixin, OracleOp3D):
    pass


def _peer_case(rank, world, ndim, shape, V, exchange, calls):
    global _MEM
    _MEM = ShmPeerMemory()
    rng = np.random.default_rng(21)
    if ndim == 2:
        angles = np.linspace(0, np.pi, V, endpoint=False)
        op = sharded.ViewShardedXRayTransform2D(shape, angles, op_factory=ScatterOp2D, exchange=exchange, peer_mem=_MEM)
        ref = OracleOp2D(shape, angles, det_count=op.ny)
        out_shape = (V, op.ny)
        part = lambda r, y: OracleOp2D(shape, angles[slice(*op.view_blocks[r])], det_count=op.ny).back_project(  # noqa: E731
            torch.from_numpy(np.ascontiguousarray(y[slice(*op.view_blocks[r])]))).numpy()
    else:
        D = (shape[0] + 3, shape[2] + 4)
        ang = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, 0.4)], 1)
        M = O.matrices_from_euler_angles(shape, D, "XY", ang).astype(np.float32)
        op = sharded.ViewShardedXRayTransform3D(shape, M, D, op_factory=ScatterOp3D, exchange=exchange, peer_mem=_MEM)
        ref = OracleOp3D(shape, M, D)
        out_shape = (V,) + D
        part = lambda r, y: OracleOp3D(shape, M[slice(*op.view_blocks[r])], D).back_project(  # noqa: E731
            torch.from_numpy(np.ascontiguousarray(y[slice(*op.view_blocks[r])]))).numpy()
    assert op.peer is not None and op.peer.mode == ("store" if exchange == "peer" else "add")
    z0, z1 = op.slab
    v0, v1 = op.views
    for it in range(calls):  # more calls than copies: every copy is reused
        y = rng.standard_normal(out_shape).astype(np.float32)
        got = op.back_project(torch.from_numpy(np.ascontiguousarray(y[v0:v1]))).numpy()
        assert got.shape == (z1 - z0,) + tuple(shape[1:])
        parts = [part(r, y)[z0:z1] if op.view_blocks[r][1] > op.view_blocks[r][0] else np.zeros_like(got)
                 for r in range(world)]
        want = parts[0].copy()
        for r in range(1, world):
            want = want + parts[r]
        if exchange == "peer":
            assert np.array_equal(got, want), it  # slots are summed in rank order: bit-exact
        else:
            assert np.allclose(got, want, rtol=1e-6, atol=1e-6), it
        assert O.rel_l2(got, ref.back_project(torch.from_numpy(y)).numpy()[z0:z1]) <= 1e-5
    op.close()
    op.close()  # idempotent


@pytest.mark.parametrize("exchange", ["peer", "peer_add"])
def test_fused_exchange_protocol_2d_three_ranks_uneven_blocks(exchange):
    _spawn(_peer_case, 2, (25, 20), 10, exchange, 5, world=3)


def test_fused_exchange_protocol_rank_without_views():
    _spawn(_peer_case, 2, (13, 12), 2, "peer", 3, world=3)  # view blocks (0,0), (0,1), (1,2)


@pytest.mark.parametrize("exchange", ["peer", "peer_add"])
def test_fused_exchange_protocol_3d_tilted(exchange):
    _spawn(_peer_case, 3, (9, 8, 7), 5, exchange, 4, world=2)


def test_fused_exchange_is_a_no_op_choice_for_a_single_rank():
    """World size 1: nothing to exchange, the operator keeps the plain back projection (no peer buffers)."""
    op = sharded.ViewShardedXRayTransform2D((12, 10), np.linspace(0, np.pi, 4, endpoint=False), op_factory=OracleOp2D,
                                            exchange="peer", rank=0, world_size=1)
    assert op.peer is None and op.exchange == "peer"
    y = np.random.default_rng(0).standard_normal((4, op.ny)).astype(np.float32)
    want = OracleOp2D((12, 10), np.linspace(0, np.pi, 4, endpoint=False), det_count=op.ny).back_project(torch.from_numpy(y))
    assert torch.equal(op.back_project(torch.from_numpy(y)), want)
    op.close()
    with pytest.raises(ValueError):
        sharded.PeerBlocks([(0, 4), (4, 8)], (3,), rank=0, world_size=3, mem=ShmPeerMemory())  # one block per rank
    with pytest.raises(ValueError):
        sharded.PeerBlocks([(0, 4)], (3,), rank=0, world_size=1, mode="carrier", mem=ShmPeerMemory())


# ---------------------------------------------------------------------------------------------
def test_block_bounds_cover_exactly():
    for n, p in ((1024, 8), (10, 3), (5, 8), (7, 7)):
        b = [sharded.block_bounds(n, p, i) for i in range(p)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(p - 1))
    with pytest.raises(ValueError):
        sharded.block_bounds(10, 2, 2)


def test_slab_plan_rows_and_ownership_single_process():
    """Partition bookkeeping without a process group (explicit rank / world_size)."""
    N, D, V = (64, 8, 8), (64, 12), 5
    M = O.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
    owned = []
    for r in range(4):
        op = sharded.SlabShardedXRayTransform3D(N, M, D, op_factory=OracleOp3D, rank=r, world_size=4)
        assert op.slab == (16 * r, 16 * r + 16) and op.rows == (16 * r, 16 * r + 16) and not op.halo_rows()
        owned.append(op.owned_rows)
    assert owned == [(0, 16), (16, 32), (32, 48), (48, 64)]
    Mh = O.matrices_from_euler_angles(N, D, "X", np.zeros((1, 1)), voxel_spacing=[0.8, 1, 1])
    ops = [sharded.SlabShardedXRayTransform3D(N, Mh, D, op_factory=OracleOp3D, rank=r, world_size=4) for r in range(4)]
    covered = sorted(sum((list(range(*o.owned_rows)) for o in ops), []))
    lo, hi = geometry.slab_row_range(Mh, 0, 64, 64)
    assert covered == list(range(lo, hi))  # each touched row owned exactly once
    assert any(o.halo_rows() for o in ops)
    with pytest.raises(ValueError):
        Mt = O.matrices_from_euler_angles(N, D, "XY", np.array([[0.3, 0.5]]))
        sharded.SlabShardedXRayTransform3D(N, Mt, D, op_factory=OracleOp3D, rank=0, world_size=2)
