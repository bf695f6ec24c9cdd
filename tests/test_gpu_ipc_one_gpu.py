"""The fused view-block exchange (``xct_adjoint_scatter`` + ``sharded.PeerBlocks``) with the REAL peer
memory on ONE GPU: several processes share ``cuda:0``, rendezvous over gloo, and map each other's
exchange buffers through CUDA IPC (which works between processes on the same device exactly as it does
across NVLink).  This is what the driver's one-GPU box can run of tests/test_gpu_multi.py: the protocol
(slot addressing, double-buffered copies, rendezvous placement, re-zeroing), the routed kernels writing
through IPC-mapped pointers, the flag rendezvous in peer memory, uneven row blocks, three ranks, a rank without
views, store and add mode -- all against the oracle.  NCCL itself needs one GPU per rank and stays in tests/test_gpu_multi.py.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, exchange, rendezvous="flags"):
    import torch
    import torch.distributed as dist

    import scico_b200 as sb
    from scico_b200 import sharded
    from oracle import xray_c as C
    from oracle import xray_np as O

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(0)
    dev = "cuda:0"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(17)  # same stream on every rank: same global arrays
        if case.startswith("view2d"):
            nx, V = {"view2d": ((200, 168), 48), "view2d_uneven": ((203, 168), 50), "view2d_idle_rank": ((96, 80), 2)}[case]
            angles = np.linspace(0, np.pi, V, endpoint=False)
            # a short rendezvous timeout: ranks time-slice ONE GPU here, a rank that cannot make progress must
            # fail the comparison below, not hang the box
            op = sharded.ViewShardedXRayTransform2D(nx, angles, exchange=exchange, rendezvous=rendezvous, peer_timeout_s=5.0)
            assert op.peer is not None and isinstance(op.peer.mem, sharded._NativePeerMemory)
            assert op.peer.rendezvous == rendezvous
            T = sb.XRayTransform2D(nx, angles).view_table
            (z0, z1), (v0, v1) = op.slab, op.views
            if case == "view2d_idle_rank":
                assert world == 3 and (rank != 0 or v1 == v0)  # rank 0 holds no view and only receives
            for it in range(5):  # both buffer copies are recycled at least twice
                y = rng.standard_normal((V, op.ny)).astype(np.float32)
                back = op.back_project(torch.as_tensor(np.ascontiguousarray(y[v0:v1]), device=dev))
                assert tuple(back.shape) == (z1 - z0, nx[1])
                assert O.rel_l2(back.cpu().numpy(), C.back_project_2d(y, T, nx)[z0:z1]) <= 1e-5, (it, rank)
            x = rng.standard_normal(nx).astype(np.float32)
            got = op.project(torch.as_tensor(x, device=dev)).cpu().numpy()
            if v1 > v0:
                assert O.rel_l2(got, C.project_2d(x, T, op.ny)[v0:v1]) <= 1e-5
            op.close()
        elif case in ("view3d_tilt", "view3d_sep"):
            if case == "view3d_tilt":  # general kernels (routed gen3d adjoint)
                N, D, V = (20, 24, 28), (30, 36), 10
                ang = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, 0.5)], 1)
                M = sb.matrices_from_euler_angles(N, D, "XY", ang)
            else:  # separable plan in view-block mode (routed walk / plane adjoint)
                N, D, V = (24, 96, 80), (24, 128), 12
                M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
            op = sharded.ViewShardedXRayTransform3D(N, M, D, exchange=exchange, rendezvous=rendezvous, peer_timeout_s=5.0)
            assert op.peer is not None
            if case == "view3d_sep":  # the routed epilogue of the walk adjoint (TMA-staged window)
                assert op.full.plan_info(0)["adj_kernel"] == 2 and op.full.plan_info(0)["adj_tma"] == 1
            else:                     # ... and of the brick adjoint
                assert op.full.plan_info(0)["adj_kernel"] == 3
            (z0, z1), (v0, v1) = op.slab, op.views
            for it in range(4):
                y = rng.standard_normal((V,) + D).astype(np.float32)
                back = op.back_project(torch.as_tensor(np.ascontiguousarray(y[v0:v1]), device=dev)).cpu().numpy()
                assert back.shape == (z1 - z0,) + N[1:]
                assert O.rel_l2(back, C.back_project_3d(y, op.matrices, N)[z0:z1]) <= 1e-5, (it, rank)
            x = rng.standard_normal(N).astype(np.float32)
            got = op.project(torch.as_tensor(x[z0:z1], device=dev)).cpu().numpy()  # all-gather of the slabs
            assert O.rel_l2(got, C.project_3d(x, op.matrices, D)[v0:v1]) <= 1e-5
            assert not op.peer.mem.timed_out()
            op.close()
        elif case == "pdhg_view3d":
            # TV-PDHG over the view-block partition of a tilted geometry: volume state in z-slabs, sinogram
            # state in view blocks, the fused peer exchange inside every back projection
            from scico_b200.optimize import TVPDHG

            N, D, V = (16, 24, 20), (26, 32), 10
            ang = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, 0.5)], 1)
            M = sb.matrices_from_euler_angles(N, D, "XY", ang)
            x_gt = np.zeros(N, np.float32)
            x_gt[3:13, 6:16, 5:14] = 1.0
            full = sb.XRayTransform3D(N, M, D)
            noise = rng.standard_normal((V,) + D).astype(np.float32)
            y = full(torch.as_tensor(x_gt, device=dev)) + 0.05 * torch.as_tensor(noise, device=dev)
            ref = TVPDHG(full, y, 0.1, 0.05, 0.05, maxiter=20, itstat=True)
            ref.solve()
            op = sharded.ViewShardedXRayTransform3D(N, M, D, exchange=exchange, rendezvous=rendezvous, peer_timeout_s=5.0)
            (z0, z1), (v0, v1) = op.slab, op.views
            S = TVPDHG(op, y[v0:v1].contiguous(), 0.1, 0.05, 0.05, maxiter=20, itstat=True)
            S.solve()
            for a, b in zip(S.history, ref.history):  # fused statistics, summed over the partition
                for key in ("objective", "prml_rsdl", "dual_rsdl"):
                    assert abs(a[key] - b[key]) <= 2e-5 * abs(b[key]) + 1e-12, (key, a, b)
            rel = (torch.linalg.vector_norm(S.x - ref.x[z0:z1]) / torch.linalg.vector_norm(ref.x[z0:z1])).item()
            assert rel <= 1e-5, (exchange, rel)
            assert abs(S.objective() - ref.objective()) <= 1e-5 * ref.objective(), exchange
            op.close()
        elif case == "pdhg_slab_stats":
            # TV-PDHG with iteration statistics over z-slabs whose detector rows overlap (|M00| = 0.7: a row collects
            # slices of two or three slabs): the statistics the iteration's kernels accumulate on every rank's OWNED
            # rows add up to those of the unsharded solver, which in turn equal the explicitly evaluated ones
            from scico_b200.optimize import TVPDHG

            N, D, V = (12, 24, 20), (8, 32), 10  # 12 slices x 0.7 = 8.4 rows: every detector row is held by a slab
            M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None],
                                              voxel_spacing=(0.7, 1.0, 1.0))
            x_gt = np.zeros(N, np.float32)
            x_gt[3:10, 6:16, 5:14] = 1.0
            full = sb.XRayTransform3D(N, M, D)
            noise = rng.standard_normal((V,) + D).astype(np.float32)
            y = full(torch.as_tensor(x_gt, device=dev)) + 0.05 * torch.as_tensor(noise, device=dev)
            ref = TVPDHG(full, y, 0.1, 0.05, 0.05, maxiter=15, itstat=True)
            ref.solve()
            op = sharded.SlabShardedXRayTransform3D(N, M, D)
            (z0, z1), (r0, r1) = op.slab, op.rows
            S = TVPDHG(op, y[:, r0:r1].contiguous(), 0.1, 0.05, 0.05, maxiter=15, itstat=True)
            S.solve()
            assert len(S.history) == 15
            for a, b in zip(S.history, ref.history):
                for key in ("objective", "prml_rsdl", "dual_rsdl"):
                    assert abs(a[key] - b[key]) <= 2e-5 * abs(b[key]) + 1e-12, (key, a, b)
            assert abs(S.history[-1]["objective"] - S.objective()) <= 1e-5 * S.objective()
        elif case in ("padmm_notebook_slabs", "padmm_notebook_views"):
            # the reference's published ct_3d_tv_padmm run (tests/golden/nb_ct_3d_tv_padmm.npz: the iteration statistics
            # its ProximalADMM printed with real JAX / XLA, see tests/test_gpu_reference_notebook.py) reproduced by the
            # z-slab PARTITION: every rank holds its slices of the 64 x 256 x 128 volume and its detector rows, the
            # statistics are summed over the ranks' owned rows -- all 1000 rows of the table, SNR and MAE
            import sys

            here = os.path.dirname(os.path.abspath(__file__))
            if here not in sys.path:
                sys.path.insert(0, here)
            import _ct3d_example as E
            from scico_b200.optimize import TVProximalADMM

            g = np.load(os.path.join(here, "golden", "nb_ct_3d_tv_padmm.npz"))
            N, M, D = E.geometry()
            x_gt = E.tangle_phantom()
            full = sb.XRayTransform3D(N, M, D)
            y = full(torch.as_tensor(x_gt, device=dev))
            if case == "padmm_notebook_views":
                # ... and by the VIEW-BLOCK partition: sinogram state in view blocks, volume state in z-slabs, every one
                # of the 1000 back projections exchanging its rows through the kernel's routed epilogue (peer memory)
                op = sharded.ViewShardedXRayTransform3D(N, M, D, exchange=exchange, rendezvous=rendezvous, peer_timeout_s=20.0)
                v0, v1 = op.views
                y_loc = y[v0:v1].contiguous()
            else:
                op = sharded.SlabShardedXRayTransform3D(N, M, D)
                r0, r1 = op.rows
                y_loc = y[:, r0:r1].contiguous()
            mu, nu = TVProximalADMM.estimate_parameters(op, alpha=E.ALPHA)
            S = TVProximalADMM(op, y_loc, E.LAM, E.RHO, mu, nu, alpha=E.ALPHA, maxiter=E.MAXITER, itstat=True)
            S.solve()
            h = S.history
            assert len(h) == E.MAXITER
            for key, tol in (("objective", 1e-3), ("prml_rsdl", 5e-3), ("dual_rsdl", 5e-3)):
                dev_ = np.abs(np.array([r[key] for r in h]) - g[key]) / g[key]
                assert dev_.max() <= tol, (key, float(dev_.max()), int(dev_.argmax()))
            parts = [torch.empty((b - a,) + tuple(N[1:])) for a, b in op.slabs]  # 64 slices: equal slabs
            dist.all_gather(parts, S.x.cpu())
            x_rec = torch.cat(parts).numpy()
            assert abs(E.snr_db(x_gt, x_rec) - float(g["snr_db"])) <= 0.02 and abs(E.mae(x_gt, x_rec) - float(g["mae"])) <= 1e-3
            if case == "padmm_notebook_views":
                assert not op.peer.mem.timed_out()
                op.close()
        else:
            raise AssertionError(case)
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        dist.destroy_process_group()


# rendezvous: "flags" = epoch words in peer memory (xct_peer_signal / xct_peer_wait, the default);
# "collective" = a one-element all-reduce of the process group
CASES = [
    ("view2d", 2, "peer", "flags"), ("view2d", 2, "peer_add", "flags"), ("view2d", 2, "peer", "collective"),
    ("view2d_uneven", 3, "peer", "flags"), ("view2d_uneven", 3, "peer_add", "collective"),
    ("view2d_idle_rank", 3, "peer", "flags"), ("view2d_idle_rank", 3, "peer_add", "flags"),
    ("view3d_tilt", 2, "peer", "flags"), ("view3d_tilt", 3, "peer", "collective"), ("view3d_tilt", 3, "peer_add", "flags"),
    ("view3d_sep", 2, "peer", "flags"), ("view3d_sep", 3, "peer_add", "flags"),
    ("pdhg_view3d", 2, "peer", "flags"),
    ("pdhg_slab_stats", 2, "nccl", "flags"), ("pdhg_slab_stats", 3, "nccl", "flags"),
    ("padmm_notebook_slabs", 2, "nccl", "flags"), ("padmm_notebook_views", 2, "peer", "flags"),
]


@pytest.mark.parametrize("case,world,exchange,rendezvous", CASES, ids=[f"{c}-w{w}-{e}-{r}" for c, w, e, r in CASES])
def test_fused_exchange_over_cuda_ipc_on_one_gpu(cuda_device, case, world, exchange, rendezvous):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), case, exchange, rendezvous), nprocs=world, join=True)
