"""Routed against plain back projection kernel on ONE GPU (no exchange): the cost of the routed epilogue.
A view block of C3 as one of two ranks holds it (4096^2, 1024 of 2048 views); the row blocks are local
buffers.  CUDA events, 3 warm-ups, 5 repetitions; writes gpurun_out/scatter_kernel_1gpu.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scico_b200 as sb

dev = "cuda:0"
n, V, parts = 4096, 2048, 2
op = sb.XRayTransform2D((n, n), np.linspace(0, np.pi, V, endpoint=False)[: V // parts], det_count=int(np.ceil(np.sqrt(2) * n)))
y = torch.rand(op.output_shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
bounds = [(n * k) // parts for k in range(parts + 1)]
blocks = [torch.zeros((b - a, n), device=dev) for a, b in zip(bounds[:-1], bounds[1:])]
ptrs = [b.data_ptr() for b in blocks]
out = torch.empty((n, n), device=dev)


def timeit(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {
    "workload": f"XRayTransform2D {n}^2, {V // parts} views (one of {parts} view blocks of C3), adjoint only, one B200",
    "plain_ms": timeit(lambda: op.back_project(y, out=out)),
    "routed_store_ms": timeit(lambda: op.back_project_scatter(y, ptrs, bounds, True)),
    "routed_add_ms": timeit(lambda: op.back_project_scatter(y, ptrs, bounds, False)),
}
op.back_project_scatter(y, ptrs, bounds, True)
res["store_vs_plain_max_abs_diff"] = float((torch.cat(blocks) - op.back_project(y)).abs().max())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/scatter_kernel_1gpu.json", "w"), indent=1)
print(json.dumps(res))
