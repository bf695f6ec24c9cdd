"""Multi-GPU (NCCL, one process per GPU) checks of scico_b200/sharded.py with the native CUDA
operators.  Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case):
    import torch
    import torch.distributed as dist

    import scico_b200 as sb
    from scico_b200 import sharded
    from oracle import xray_c as C
    from oracle import xray_np as O

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        rng = np.random.default_rng(3)
        if case == "slab":
            N, D, V = (32, 48, 40), (32, 64), 12
            M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
            x = rng.standard_normal(N).astype(np.float32)
            y = rng.standard_normal((V,) + D).astype(np.float32)
            op = sharded.SlabShardedXRayTransform3D(N, M, D)
            (z0, z1), (r0, r1) = op.slab, op.rows
            got = op.project(torch.as_tensor(x[z0:z1], device=dev)).cpu().numpy()
            assert O.rel_l2(got, C.project_3d(x, op.matrices, D)[:, r0:r1]) <= 1e-5
            back = op.back_project(torch.as_tensor(np.ascontiguousarray(y[:, r0:r1]), device=dev)).cpu().numpy()
            assert O.rel_l2(back, C.back_project_3d(y, op.matrices, N)[z0:z1]) <= 1e-5
        elif case == "slab_halo":
            N, D, V = (30, 40, 36), (32, 52), 6
            M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None],
                                              voxel_spacing=[0.8, 1.0, 1.0])
            x = rng.standard_normal(N).astype(np.float32)
            op = sharded.SlabShardedXRayTransform3D(N, M, D)
            (z0, z1), (r0, r1) = op.slab, op.rows
            got = op.project(torch.as_tensor(x[z0:z1], device=dev)).cpu().numpy()
            assert O.rel_l2(got, C.project_3d(x, op.matrices, D)[:, r0:r1]) <= 1e-5
        elif case == "view3d":
            N, D, V = (20, 24, 28), (30, 36), 10
            ang = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, 0.5)], 1)
            M = sb.matrices_from_euler_angles(N, D, "XY", ang)
            x = rng.standard_normal(N).astype(np.float32)
            y = rng.standard_normal((V,) + D).astype(np.float32)
            op = sharded.ViewShardedXRayTransform3D(N, M, D)
            (z0, z1), (v0, v1) = op.slab, op.views
            got = op.project(torch.as_tensor(x[z0:z1], device=dev)).cpu().numpy()
            assert O.rel_l2(got, C.project_3d(x, op.matrices, D)[v0:v1]) <= 1e-5
            back = op.back_project(torch.as_tensor(np.ascontiguousarray(y[v0:v1]), device=dev)).cpu().numpy()
            assert O.rel_l2(back, C.back_project_3d(y, op.matrices, N)[z0:z1]) <= 1e-5
        elif case == "view2d":
            nx, V = (96, 80), 30
            angles = np.linspace(0, np.pi, V, endpoint=False)
            op = sharded.ViewShardedXRayTransform2D(nx, angles)
            full = sb.XRayTransform2D(nx, angles)
            x = rng.standard_normal(nx).astype(np.float32)
            y = rng.standard_normal((V, op.ny)).astype(np.float32)
            T = full.view_table
            (z0, z1), (v0, v1) = op.slab, op.views
            got = op.project(torch.as_tensor(x, device=dev)).cpu().numpy()
            assert O.rel_l2(got, C.project_2d(x, T, op.ny)[v0:v1]) <= 1e-5
            back = op.back_project(torch.as_tensor(np.ascontiguousarray(y[v0:v1]), device=dev)).cpu().numpy()
            assert O.rel_l2(back, C.back_project_2d(y, T, nx)[z0:z1]) <= 1e-5
        elif case in ("view2d_peer", "view2d_peer_add"):
            # back projection fused with the exchange: rows added into their owners over NVLink peer memory
            nx, V = (200, 168), 48
            angles = np.linspace(0, np.pi, V, endpoint=False)
            op = sharded.ViewShardedXRayTransform2D(nx, angles, exchange="peer" if case == "view2d_peer" else "peer_add")
            nccl = sharded.ViewShardedXRayTransform2D(nx, angles)
            T = sb.XRayTransform2D(nx, angles).view_table
            (z0, z1), (v0, v1) = op.slab, op.views
            assert op.peer is not None
            for it in range(4):  # both block copies are reused
                y = rng.standard_normal((V, op.ny)).astype(np.float32)
                yv = torch.as_tensor(np.ascontiguousarray(y[v0:v1]), device=dev)
                back = op.back_project(yv)
                assert tuple(back.shape) == (z1 - z0, nx[1])
                assert O.rel_l2(back.cpu().numpy(), C.back_project_2d(y, T, nx)[z0:z1]) <= 1e-5, it
                ref = nccl.back_project(yv)
                assert (torch.linalg.vector_norm(back - ref) / torch.linalg.vector_norm(ref)).item() <= 1e-6
            op.close()
        elif case in ("view3d_peer", "view3d_peer_add"):
            N, D, V = (20, 24, 28), (30, 36), 10
            ang = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, 0.5)], 1)
            M = sb.matrices_from_euler_angles(N, D, "XY", ang)
            op = sharded.ViewShardedXRayTransform3D(N, M, D, exchange="peer" if case == "view3d_peer" else "peer_add")
            (z0, z1), (v0, v1) = op.slab, op.views
            for it in range(3):
                y = rng.standard_normal((V,) + D).astype(np.float32)
                back = op.back_project(torch.as_tensor(np.ascontiguousarray(y[v0:v1]), device=dev)).cpu().numpy()
                assert O.rel_l2(back, C.back_project_3d(y, op.matrices, N)[z0:z1]) <= 1e-5, it
            x = rng.standard_normal(N).astype(np.float32)
            got = op.project(torch.as_tensor(x[z0:z1], device=dev)).cpu().numpy()
            assert O.rel_l2(got, C.project_3d(x, op.matrices, D)[v0:v1]) <= 1e-5
            op.close()
        elif case == "pdhg_slab":
            from scico_b200.optimize import TVPDHG

            N, D, V = (16, 24, 20), (16, 32), 10
            M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
            x_gt = np.zeros(N, np.float32)
            x_gt[3:13, 6:16, 5:14] = 1.0
            full = sb.XRayTransform3D(N, M, D)
            y = full(torch.as_tensor(x_gt, device=dev)) + 0.05 * torch.randn((V,) + D, device=dev,
                                                                              generator=torch.Generator(device=dev).manual_seed(1))
            dist.broadcast(y, src=0)  # same data on both ranks
            ref = TVPDHG(full, y, 0.1, 0.05, 0.05, maxiter=20)
            ref.solve()
            op = sharded.SlabShardedXRayTransform3D(N, M, D)
            (z0, z1), (r0, r1) = op.slab, op.rows
            S = TVPDHG(op, y[:, r0:r1].contiguous(), 0.1, 0.05, 0.05, maxiter=20)
            S.solve()
            rel = (torch.linalg.vector_norm(S.x - ref.x[z0:z1]) / torch.linalg.vector_norm(ref.x[z0:z1])).item()
            assert rel <= 1e-5, rel
            assert abs(S.objective() - ref.objective()) <= 1e-5 * ref.objective()
        elif case == "pdhg_view3d":
            # TV-PDHG over the view-block partition of a tilted geometry: volume state in z-slabs, sinogram
            # state in view blocks, fused peer exchange in the back projection
            from scico_b200.optimize import TVPDHG

            N, D, V = (16, 24, 20), (26, 32), 10
            ang = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, 0.5)], 1)
            M = sb.matrices_from_euler_angles(N, D, "XY", ang)
            x_gt = np.zeros(N, np.float32)
            x_gt[3:13, 6:16, 5:14] = 1.0
            full = sb.XRayTransform3D(N, M, D)
            y = full(torch.as_tensor(x_gt, device=dev)) + 0.05 * torch.randn((V,) + D, device=dev,
                                                                              generator=torch.Generator(device=dev).manual_seed(1))
            dist.broadcast(y, src=0)
            ref = TVPDHG(full, y, 0.1, 0.05, 0.05, maxiter=20)
            ref.solve()
            for exchange in ("nccl", "peer"):
                op = sharded.ViewShardedXRayTransform3D(N, M, D, exchange=exchange)
                (z0, z1), (v0, v1) = op.slab, op.views
                S = TVPDHG(op, y[v0:v1].contiguous(), 0.1, 0.05, 0.05, maxiter=20)
                S.solve()
                rel = (torch.linalg.vector_norm(S.x - ref.x[z0:z1]) / torch.linalg.vector_norm(ref.x[z0:z1])).item()
                assert rel <= 1e-5, (exchange, rel)
                assert abs(S.objective() - ref.objective()) <= 1e-5 * ref.objective(), exchange
                op.close()
        elif case in ("admm_slab", "ladmm_slab", "padmm_slab"):
            from scico_b200.optimize import TVADMM, TVLinearizedADMM, TVProximalADMM

            N, D, V = (16, 24, 20), (16, 32), 10
            M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
            x_gt = np.zeros(N, np.float32)
            x_gt[3:13, 6:16, 5:14] = 1.0
            full = sb.XRayTransform3D(N, M, D)
            y = full(torch.as_tensor(x_gt, device=dev)) + 0.05 * torch.randn((V,) + D, device=dev,
                                                                              generator=torch.Generator(device=dev).manual_seed(1))
            dist.broadcast(y, src=0)
            op = sharded.SlabShardedXRayTransform3D(N, M, D)
            (z0, z1), (r0, r1) = op.slab, op.rows
            yl = y[:, r0:r1].contiguous()
            if case == "admm_slab":
                mk = lambda A, yy: TVADMM(A, yy, 0.5, 5.0, maxiter=5, cg_tol=1e-30, cg_maxiter=6)  # noqa: E731
            elif case == "ladmm_slab":
                mk = lambda A, yy: TVLinearizedADMM(A, yy, 0.1, 1.0 / 130.0, 1.0, maxiter=20)  # noqa: E731
            else:
                mk = lambda A, yy: TVProximalADMM(A, yy, 0.1, 0.05, 400.0, 1.01, alpha=4.0, maxiter=20, itstat=True)  # noqa: E731
            ref, S = mk(full, y), mk(op, yl)
            ref.solve()
            S.solve()
            rel = (torch.linalg.vector_norm(S.x - ref.x[z0:z1]) / torch.linalg.vector_norm(ref.x[z0:z1])).item()
            assert rel <= 1e-5, rel
            if case == "padmm_slab":
                assert abs(S.history[-1]["objective"] - ref.history[-1]["objective"]) <= 1e-5 * ref.history[-1]["objective"]
                assert abs(S.history[-1]["prml_rsdl"] - ref.history[-1]["prml_rsdl"]) <= 1e-4 * ref.history[-1]["prml_rsdl"]
        elif case == "padmm_notebook_slabs":
            # the reference's published ct_3d_tv_padmm run reproduced over two z-slabs on two GPUs (NCCL): all 1000 rows
            # of the notebook's iteration-statistics table (see tests/test_gpu_reference_notebook.py)
            import sys

            here = os.path.dirname(os.path.abspath(__file__))
            if here not in sys.path:
                sys.path.insert(0, here)
            import _ct3d_example as E
            from scico_b200.optimize import TVProximalADMM

            g = np.load(os.path.join(here, "golden", "nb_ct_3d_tv_padmm.npz"))
            N, M, D = E.geometry()
            x_gt = E.tangle_phantom()
            y = sb.XRayTransform3D(N, M, D)(torch.as_tensor(x_gt, device=dev))
            op = sharded.SlabShardedXRayTransform3D(N, M, D)
            (z0, z1), (r0, r1) = op.slab, op.rows
            mu, nu = TVProximalADMM.estimate_parameters(op, alpha=E.ALPHA)
            S = TVProximalADMM(op, y[:, r0:r1].contiguous(), E.LAM, E.RHO, mu, nu, alpha=E.ALPHA, maxiter=E.MAXITER, itstat=True)
            S.solve()
            h = S.history
            assert len(h) == E.MAXITER
            for key, tol in (("objective", 1e-3), ("prml_rsdl", 5e-3), ("dual_rsdl", 5e-3)):
                d = np.abs(np.array([r[key] for r in h]) - g[key]) / g[key]
                assert d.max() <= tol, (key, float(d.max()), int(d.argmax()))
            assert abs(E.snr_db(x_gt[z0:z1], S.x.cpu().numpy()) - float(g["snr_db"])) <= 1.0  # this rank's half of the volume
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["slab", "slab_halo", "view3d", "view2d", "view2d_peer", "view3d_peer", "view2d_peer_add",
                                  "view3d_peer_add", "pdhg_slab", "pdhg_view3d", "admm_slab", "ladmm_slab", "padmm_slab",
                                  "padmm_notebook_slabs"])
def test_sharded_operators_nccl(case):
    import torch
    import torch.multiprocessing as mp

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), case), nprocs=2, join=True)
