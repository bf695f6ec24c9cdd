"""xct_adjoint_scatter: the back projection whose epilogue adds every result row into the row block of
its owner (view-block sharding fused with its exchange step).

On ONE GPU the row blocks are separate local buffers: this checks the routing of every routed kernel
(2D plane / 2D general / 3D plane / 3D brick / 3D thread-per-voxel) against the plain back projection and the oracle, the
error behaviour of the entry point, and the PeerBlocks protocol (double-buffered blocks, copy-out,
re-zeroing) with world size 1.  The cross-GPU part (CUDA IPC over NVLink) is tests/test_gpu_multi.py.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scico_b200 as sb
from scico_b200 import _lib, sharded
from oracle import xray_c as C
from oracle import xray_np as O


def _x_mats(N, D, V, **kw):
    return sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None], **kw)


def _tilt_mats(N, D, V):
    ang = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, np.deg2rad(74.0))], 1)
    return sb.matrices_from_euler_angles(N, D, "XY", ang)


def _scatter(torch, dev, op, y, bounds, store=False):
    """Run back_project_scatter into len(bounds)-1 separate blocks (zeroed for the add mode, poisoned for the
    store mode, which must overwrite every element); returns them concatenated."""
    inner = tuple(op.input_shape[1:])
    blocks = [torch.full((b - a,) + inner, 7.0 if store else 0.0, dtype=torch.float32, device=dev)
              for a, b in zip(bounds[:-1], bounds[1:])]
    ptrs = [blk.data_ptr() if blk.numel() else 0 for blk in blocks]
    op.back_project_scatter(y, ptrs, bounds, store)
    return torch.cat(blocks, dim=0)


CASES = {
    # name: (operator factory, expected adjoint kernel family, row boundaries)
    "2d_plane": (lambda: sb.XRayTransform2D((96, 80), np.linspace(0, np.pi, 30, endpoint=False)), 1, [0, 31, 64, 96]),
    "2d_plane_many_views": (lambda: sb.XRayTransform2D((300, 280), np.linspace(0, np.pi, 64, endpoint=False)), 1,
                            [0, 37, 75, 112, 150, 187, 225, 262, 300]),
    "2d_general": (lambda: sb.XRayTransform2D((40, 36), np.linspace(0, np.pi, 9, endpoint=False),
                                              _flags=_lib.FLAG_FORCE_GENERAL), 0, [0, 13, 13, 40]),
    "3d_plane": (lambda: sb.XRayTransform3D((17, 18, 19), _x_mats((17, 18, 19), (20, 21), 5), (20, 21)), 1, [0, 8, 17]),
    "3d_sep_walk_plan": (lambda: sb.XRayTransform3D((24, 96, 80), _x_mats((24, 96, 80), (24, 128), 12), (24, 128)), 2,
                         [0, 6, 12, 18, 24]),
    # the scalar-tap walk adjoint (the default plan above routes through the slice-interleaved kernel)
    "3d_sep_walk_plan_scalar_taps": (lambda: sb.XRayTransform3D((24, 96, 80), _x_mats((24, 96, 80), (24, 128), 12), (24, 128),
                                                                 _flags=_lib.FLAG_NO_ADJ_VEC), 2, [0, 6, 12, 18, 24]),
    "3d_general": (lambda: sb.XRayTransform3D((17, 18, 19), _tilt_mats((17, 18, 19), (20, 21), 5), (20, 21)), 3,
                   [0, 5, 11, 17]),  # brick adjoint, cp.async-staged window (odd detector width)
    "3d_general_tma": (lambda: sb.XRayTransform3D((21, 30, 37), _tilt_mats((21, 30, 37), (44, 48), 7), (44, 48)), 3,
                       [0, 3, 9, 9, 16, 21]),  # brick adjoint, TMA-staged window; blocks cut bricks, one block empty
    "3d_general_thread_per_voxel": (lambda: sb.XRayTransform3D((17, 18, 19), _tilt_mats((17, 18, 19), (20, 21), 5), (20, 21),
                                                                _flags=_lib.FLAG_NO_BRICK), 0, [0, 5, 11, 17]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_scatter_matches_back_projection(cuda_device, name):
    import torch

    make, kernel, bounds = CASES[name]
    op = make()
    assert op.plan_info()["adj_kernel"] == kernel
    rng = np.random.default_rng(11)
    y_np = rng.standard_normal(op.output_shape).astype(np.float32)
    y = torch.as_tensor(y_np, device=cuda_device)
    want = op.adj(y)
    got = _scatter(torch, cuda_device, op, y, bounds)
    assert got.shape == want.shape
    rel = (torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want)).item()
    assert rel <= 1e-6, rel
    # and against the oracle, at the path's tolerance
    if len(op.input_shape) == 2:
        ref = C.back_project_2d(y_np, O.view_table_2d(op.angles, op.x0, op.dx, op.y0), op.input_shape)
    else:
        ref = C.back_project_3d(y_np, op.matrices, op.input_shape)
    assert O.rel_l2(got.cpu().numpy(), ref) <= 1e-5
    # store mode: plain stores, every element written exactly once (the blocks start out poisoned)
    got_s = _scatter(torch, cuda_device, op, y, bounds, store=True)
    rel_s = (torch.linalg.vector_norm(got_s - want) / torch.linalg.vector_norm(want)).item()
    assert rel_s <= 1e-6, rel_s
    # the blocks are ADDED to: a second application doubles them
    inner = tuple(op.input_shape[1:])
    blk = torch.zeros(tuple(op.input_shape), dtype=torch.float32, device=cuda_device)
    for _ in range(2):
        op.back_project_scatter(y, [blk.data_ptr()], [0, op.input_shape[0]])
    rel2 = (torch.linalg.vector_norm(blk - 2 * want) / torch.linalg.vector_norm(want)).item()
    assert rel2 <= 1e-6, (rel2, inner)


def test_scatter_rejects_bad_routes(cuda_device):
    import torch

    op = sb.XRayTransform2D((32, 24), np.linspace(0, np.pi, 6, endpoint=False))
    y = torch.zeros(op.output_shape, dtype=torch.float32, device=cuda_device)
    buf = torch.zeros((32, 24), dtype=torch.float32, device=cuda_device)
    p = buf.data_ptr()
    with pytest.raises(_lib.XctError):  # does not cover all rows
        op.back_project_scatter(y, [p], [0, 31])
    with pytest.raises(_lib.XctError):  # does not start at row 0
        op.back_project_scatter(y, [p], [1, 32])
    with pytest.raises(_lib.XctError):  # decreasing boundaries
        op.back_project_scatter(y, [p, p, p], [0, 20, 10, 32])
    with pytest.raises(_lib.XctError):  # null pointer for a non-empty block
        op.back_project_scatter(y, [p, 0], [0, 16, 32])
    with pytest.raises(ValueError):  # more blocks than the route holds
        op.back_project_scatter(y, [p] * 17, list(range(17)) + [32])
    with pytest.raises(ValueError):
        op.back_project_scatter(torch.zeros((3, 3), device=cuda_device), [p], [0, 32])
    with pytest.raises(ValueError):  # host arrays have no routed path
        op.back_project_scatter(np.zeros(op.output_shape, np.float32), [p], [0, 32])
    op.back_project_scatter(y, [p, 0, p], [0, 16, 16, 32])  # an empty block may be null


def test_sum_slots(cuda_device):
    import torch

    L = _lib.lib()
    for n, pitch, nslots in [(1024, 1024, 3), (1001, 1003, 5), (8, 8, 1), (4100, 4200, 16)]:
        g = torch.Generator(device=cuda_device).manual_seed(n)
        slots = torch.randn((nslots, pitch), device=cuda_device, generator=g)
        dst = torch.full((n,), 9.0, device=cuda_device)
        _lib.check(L.xct_sum_slots(0, dst.data_ptr(), slots.data_ptr(), nslots, n, pitch,
                                   torch.cuda.current_stream().cuda_stream))
        want = slots[0, :n].clone()
        for s_ in range(1, nslots):  # same association order: bit-exact
            want = want + slots[s_, :n]
        assert torch.equal(dst, want), (n, pitch, nslots)
    with pytest.raises(_lib.XctError):
        _lib.check(L.xct_sum_slots(0, dst.data_ptr(), slots.data_ptr(), 0, 8, 8, None))


@pytest.mark.parametrize("mode", ["store", "add"])
def test_peer_blocks_protocol_single_rank(cuda_device, mode):
    """World size 1: allocation, double-buffered exchange, slot sum / copy-out and re-zeroing of the blocks."""
    import torch

    nx, V = (64, 48), 20
    angles = np.linspace(0, np.pi, V, endpoint=False)
    op = sb.XRayTransform2D(nx, angles)
    pb = sharded.PeerBlocks([(0, nx[0])], nx[1:], rank=0, world_size=1, mode=mode)
    try:
        assert pb.local_shape == nx and len(pb.ptrs) == 2 and pb.ptrs[0] != pb.ptrs[1]
        rng = np.random.default_rng(5)
        for it in range(5):  # both copies are used more than once: each must come back zeroed
            y = torch.as_tensor(rng.standard_normal(op.output_shape).astype(np.float32), device=cuda_device)
            out = torch.empty(nx, dtype=torch.float32, device=cuda_device)
            pb.exchange(lambda ptrs, rb, st: op.back_project_scatter(y, ptrs, rb, st), out)
            want = op.adj(y)
            rel = (torch.linalg.vector_norm(out - want) / torch.linalg.vector_norm(want)).item()
            assert rel <= 1e-6, (it, rel)
        with pytest.raises(ValueError):
            pb.exchange(lambda ptrs, rb, st: None, torch.empty((3, 3), device=cuda_device))
    finally:
        pb.close()
    pb.close()  # idempotent
