"""View-block back projection: NCCL exchange against the fused peer-memory exchange.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29521 tools/bench_peer_exchange.py

* C3 (BASELINE.json configs[2]): 2D 4096^2, 2048 views, view blocks.  "nccl": plane adjoint kernel writes a
  partial image (64 MB), one NCCL reduce_scatter sums the row blocks into their owners.  "peer": ONE kernel
  (xct_adjoint_scatter) whose epilogue adds every image row into its owner's block through NVLink peer
  memory (plain stores into per-rank slots that the owner sums, or atomics), one one-element all-reduce as
  the rendezvous.
* tilted 3D (general matrices, 256^3 x 64 views at 74 degrees): "nccl": one kernel + one NCCL reduce per
  destination slab; "peer": one routed kernel.

Both variants are checked against each other before they are timed.  CUDA events bracketed by barriers,
3 warm-ups, max over ranks; rank 0 prints one JSON object and writes gpurun_out/peer_exchange_<N>gpu.json."""
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


def timeit(fn, reps=5):
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


def rel(a, b):
    return reduce_max((torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item())


out = {"n_gpus": world}
g = torch.Generator(device=dev).manual_seed(1234)
small = "--small" in sys.argv

# ---- C3: 2D view blocks
n, V = (1024, 512) if small else (4096, 2048)
angles = np.linspace(0, np.pi, V, endpoint=False)
A_nccl = sharded.ViewShardedXRayTransform2D((n, n), angles)
A_peer = sharded.ViewShardedXRayTransform2D((n, n), angles, exchange="peer")
A_add = sharded.ViewShardedXRayTransform2D((n, n), angles, exchange="peer_add")
y = torch.rand(A_nccl.local_output_shape, device=dev, generator=g)
r = rel(A_peer.back_project(y), A_nccl.back_project(y))
r_add = rel(A_add.back_project(y), A_nccl.back_project(y))
t_nccl = timeit(lambda: A_nccl.back_project(y))
t_peer = timeit(lambda: A_peer.back_project(y))
t_add = timeit(lambda: A_add.back_project(y))
t_kern = timeit(lambda: A_nccl.local.back_project(y))
out[f"C3 2D {n}^2 x {V} views, view blocks, adjoint"] = {
    "views_per_rank": A_nccl.views[1] - A_nccl.views[0], "adj_ms_nccl_reduce_scatter": t_nccl,
    "adj_ms_peer_fused_store_slots": t_peer, "adj_ms_peer_fused_atomics": t_add,
    "adj_ms_kernel_only_no_exchange": t_kern, "peer_vs_nccl_rel_l2": r, "peer_add_vs_nccl_rel_l2": r_add}
A_peer.close()
A_add.close()
del A_nccl, A_peer, A_add, y
torch.cuda.empty_cache()

# ---- tilted 3D: general kernels
n, V = (96, 32) if small else (256, 64)
angs = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, np.deg2rad(74.0))], 1)
D = (n + 64, n + 64)
M = sb.matrices_from_euler_angles((n,) * 3, D, "XY", angs)
A_nccl = sharded.ViewShardedXRayTransform3D((n,) * 3, M, D)
A_peer = sharded.ViewShardedXRayTransform3D((n,) * 3, M, D, exchange="peer")
A_add = sharded.ViewShardedXRayTransform3D((n,) * 3, M, D, exchange="peer_add")
ys = torch.rand(A_nccl.local_output_shape, device=dev, generator=g)
r = rel(A_peer.back_project(ys), A_nccl.back_project(ys))
r_add = rel(A_add.back_project(ys), A_nccl.back_project(ys))
t_nccl = timeit(lambda: A_nccl.back_project(ys))
t_peer = timeit(lambda: A_peer.back_project(ys))
t_add = timeit(lambda: A_add.back_project(ys))
out[f"3D {n}^3 x {V} views, XY tilt 74 deg, view blocks, adjoint"] = {
    "adj_ms_nccl_per_slab_reduce": t_nccl, "adj_ms_peer_fused_store_slots": t_peer, "adj_ms_peer_fused_atomics": t_add,
    "peer_vs_nccl_rel_l2": r, "peer_add_vs_nccl_rel_l2": r_add}
A_peer.close()
A_add.close()

if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/peer_exchange_{world}gpu.json", "w"), indent=1)
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
