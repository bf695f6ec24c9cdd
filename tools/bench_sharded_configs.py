"""Multi-GPU timings of the BASELINE.json configurations that partition by VIEW BLOCKS (the one
exchange step of the path) and of C4's z-slabs.  Run under torchrun on N GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29520 tools/bench_sharded_configs.py

* C3 (configs[2]): 2D XRayTransform2D 4096^2, 2048 views, 5793 bins; rank r holds views [v0, v1); the
  image stays replicated (64 MB).  project = local kernels; back_project = local kernels + the partial
  images sum-reduced row block by row block into their owners (NCCL reduce, the reduce-scatter of the
  configuration's name) or all-reduced when the solver keeps x replicated.
* tilted 3D (general matrices, 256^3 x 64 views at 74 degrees): all-gather of the slab-sharded volume +
  local kernels / local kernels + per-slab NCCL reduce.
* C4 (configs[3]): 3D 512^3, 720 views, det 512^2 in z-slabs, no data-path collective.

CUDA events bracketed by barriers, 3 warm-ups, max over ranks; rank 0 prints one JSON object and
writes gpurun_out/sharded_configs_<N>gpu.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import scico_b200 as sb
from scico_b200 import sharded

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    os.environ.pop("NCCL_DEBUG", None)
    dist.init_process_group("nccl", device_id=torch.device(dev))


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def reduce_max(v):
    if world == 1:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timeit(fn, reps=3):
    for _ in range(3):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    return reduce_max(e0.elapsed_time(e1) / reps)


out = {"n_gpus": world}
g = torch.Generator(device=dev).manual_seed(1234)  # same seed on every rank: replicated inputs agree

# ---- C3: view-block 2D
n, V = 4096, 2048
A = sharded.ViewShardedXRayTransform2D((n, n), np.linspace(0, np.pi, V, endpoint=False))
x = torch.rand((n, n), device=dev, generator=g)
y = A.project(x)
f_ms = timeit(lambda: A.project(x))
a_scatter = timeit(lambda: A.back_project(y, scatter=True))
a_allred = timeit(lambda: A.back_project(y, scatter=False))
a_local = timeit(lambda: A.local.back_project(y))
a_peer = None
if world > 1:  # back projection fused with the exchange over NVLink peer memory (xct_adjoint_scatter)
    P = sharded.ViewShardedXRayTransform2D((n, n), np.linspace(0, np.pi, V, endpoint=False), exchange="peer")
    a_peer = timeit(lambda: P.back_project(y))
    P.close()
    del P
upd = float(n) * n * V
out["C3 2D 4096^2 x 2048 views, view blocks"] = {
    "views_per_rank": A.views[1] - A.views[0], "fwd_ms": f_ms, "adj_ms_reduce_scatter": a_scatter,
    "adj_ms_all_reduce": a_allred, "adj_ms_kernels_only": a_local, "adj_ms_fused_peer_exchange": a_peer,
    "pair_updates_per_s": 2 * upd / (f_ms + a_scatter) * 1e3,
    "exchange": "partial images (64 MB per rank): one NCCL reduce_scatter over equal row blocks (per-block reduce when the rows do not divide evenly)"}
del A, x, y
torch.cuda.empty_cache()

# ---- tilted 3D: view-block with all-gather / per-slab reduce
n, V = 256, 64
angs = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, np.deg2rad(74.0))], 1)
D = (n + 64, n + 64)
M = sb.matrices_from_euler_angles((n,) * 3, D, "XY", angs)
A = sharded.ViewShardedXRayTransform3D((n,) * 3, M, D)
xs = torch.rand(A.local_input_shape, device=dev, generator=g)
ys = A.project(xs)
f_ms = timeit(lambda: A.project(xs))
a_ms = timeit(lambda: A.back_project(ys))
a_peer = None
if world > 1:
    P = sharded.ViewShardedXRayTransform3D((n,) * 3, M, D, exchange="peer")
    a_peer = timeit(lambda: P.back_project(ys))
    P.close()
    del P
upd = float(n) ** 3 * V
out["3D 256^3 x 64 views, XY tilt 74 deg, view blocks (general kernels)"] = {
    "fwd_ms": f_ms, "adj_ms": a_ms, "adj_ms_fused_peer_exchange": a_peer, "pair_updates_per_s": 2 * upd / (f_ms + a_ms) * 1e3,
    "exchange": "forward: all-gather of the slab-sharded volume; adjoint: per-slab NCCL reduce overlapped with the next slab's kernels"}
del A, xs, ys
torch.cuda.empty_cache()

# ---- C4: z-slabs
n, V = 512, 720
M = sb.matrices_from_euler_angles((n,) * 3, (n, n), "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
SA = sharded.SlabShardedXRayTransform3D((n,) * 3, M, (n, n), rank=rank, world_size=world)
A = SA.local
xs = torch.rand(A.input_shape, device=dev, generator=g)
ys = A(xs)
f_ms = timeit(lambda: A(xs))
a_ms = timeit(lambda: A.adj(ys))
upd = float(n) ** 3 * V
out["C4 3D 512^3 x 720 views, det 512^2, z-slabs"] = {
    "fwd_ms": f_ms, "adj_ms": a_ms, "pair_updates_per_s": 2 * upd / (f_ms + a_ms) * 1e3, "exchange": "none"}

if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/sharded_configs_{world}gpu.json", "w"), indent=1)
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
