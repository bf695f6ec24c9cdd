"""What bounds the end-to-end (host arrays in, host arrays out) step when N ranks share one host?

Run under torchrun on N GPUs.  Per rank, for the z-slab of the 1024^3 x 1024-view operator this rank owns:
  * raw copy rates with all ranks copying at once: H2D alone, D2H alone, both directions together (pinned memory,
    one cudaMemcpyAsync of the slab's volume / sinogram per direction);
  * the device-only step (forward + adjoint, inputs resident);
  * the end-to-end step through project / back_project on pinned host arrays for several slice-chunk sizes of the
    host pipeline (XCT_HOST_CHUNK_SLICES), dependent pair and overlapped pair.
Rank 0 prints one JSON object (max over ranks for times, min for rates)."""
import json
import os
import sys
import time

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
    dist.init_process_group("nccl", device_id=torch.device(dev))


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def red(v, op):
    if world == 1:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=op)
    return float(t.item())


n, V = 1024, 1024
M = sb.matrices_from_euler_angles((n,) * 3, (n, n), "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
out = {"n_gpus": world, "cpus_allowed": len(os.sched_getaffinity(0))}
x = torch.randn((n // world, n, n), device=dev)
xh = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
xh.copy_(x)
xo = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
x2 = torch.empty_like(x)
nbytes = x.numel() * 4
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def rate(fn, reps=3):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    barrier()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def h2d():
    with torch.cuda.stream(s1):
        x2.copy_(xh, non_blocking=True)
    s1.synchronize()


def d2h():
    with torch.cuda.stream(s2):
        xo.copy_(x, non_blocking=True)
    s2.synchronize()


def both():
    with torch.cuda.stream(s1):
        x2.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2):
        xo.copy_(x, non_blocking=True)
    s1.synchronize()
    s2.synchronize()


out["h2d_alone_gb_s_per_gpu"] = red(rate(h2d), dist.ReduceOp.MIN if world > 1 else None)
out["d2h_alone_gb_s_per_gpu"] = red(rate(d2h), dist.ReduceOp.MIN if world > 1 else None)
out["both_directions_gb_s_per_gpu_per_direction"] = red(rate(both), dist.ReduceOp.MIN if world > 1 else None)
del x2

SA = sharded.SlabShardedXRayTransform3D((n,) * 3, M, (n, n), rank=rank, world_size=world)
A = SA.local
y = A(x)
sh = torch.empty(y.shape, dtype=torch.float32, pin_memory=True)
sh.copy_(y)
so = torch.empty(y.shape, dtype=torch.float32, pin_memory=True)
for _ in range(2):
    A(x); A.adj(y)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    A(x); A.adj(y)
e1.record()
barrier()
out["device_step_ms"] = red(e0.elapsed_time(e1) / 3, dist.ReduceOp.MAX if world > 1 else None)
out["bytes_per_direction_per_gpu_per_step"] = 2 * nbytes

xh_np, sh_np, so_np, xo_np = xh.numpy(), sh.numpy(), so.numpy(), xo.numpy()
out["e2e"] = {}
for chunk in (0, 8, 16, 32, 64):
    if chunk:
        os.environ["XCT_HOST_CHUNK_SLICES"] = str(chunk)
    else:
        os.environ.pop("XCT_HOST_CHUNK_SLICES", None)
    if chunk > n // world:
        continue
    A.project(xh_np, out=so_np); A.back_project(sh_np, out=xo_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        A.project(xh_np, out=so_np)
        A.back_project(so_np, out=xo_np)
    barrier()
    dep = red((time.perf_counter() - t0) / 3 * 1e3, dist.ReduceOp.MAX if world > 1 else None)
    t0 = time.perf_counter()
    for _ in range(3):
        A.project(xh_np, out=so_np, wait=False)
        A.back_project(sh_np, out=xo_np, wait=False)
        A.host_wait()
    barrier()
    ov = red((time.perf_counter() - t0) / 3 * 1e3, dist.ReduceOp.MAX if world > 1 else None)
    # forward alone / adjoint alone
    t0 = time.perf_counter()
    for _ in range(3):
        A.project(xh_np, out=so_np)
    barrier()
    f = red((time.perf_counter() - t0) / 3 * 1e3, dist.ReduceOp.MAX if world > 1 else None)
    t0 = time.perf_counter()
    for _ in range(3):
        A.back_project(sh_np, out=xo_np)
    barrier()
    a = red((time.perf_counter() - t0) / 3 * 1e3, dist.ReduceOp.MAX if world > 1 else None)
    out["e2e"]["default" if not chunk else f"chunk_{chunk}"] = {"dependent_pair_ms": dep, "overlapped_pair_ms": ov, "forward_ms": f, "adjoint_ms": a}
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/host_pipeline_{world}gpu.json", "w"), indent=1)
    print(json.dumps(out, indent=1))
if world > 1:
    dist.destroy_process_group()
