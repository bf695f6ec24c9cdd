"""Developer A/B timing of kernel generations on the B200 (run under gpurun).

    python tools/kernel_ab.py [--size 1024] [--views 256] [--slices 128]

Times xct_forward / xct_adjoint for the default plan and for the plan with XCT_FLAG_NO_WALK
(first-generation plane kernels) on a z-slab of the headline workload, and checks they agree.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scico_b200 as sb
from scico_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--views", type=int, default=256)
ap.add_argument("--slices", type=int, default=128)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

n, V, S = args.size, args.views, args.slices
dev = "cuda:0"
M = sb.matrices_from_euler_angles((n, n, n), (n, n), "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
z0 = (n - S) // 2
kw = dict(slice_offset=z0, det_row_offset=z0, det_rows_total=n)
ops = {
    "default": sb.XRayTransform3D((S, n, n), M, (S, n), **kw),
    "no_joint": sb.XRayTransform3D((S, n, n), M, (S, n), _flags=_lib.FLAG_NO_JOINT, **kw),  # one-column walk forward
    "no_tma": sb.XRayTransform3D((S, n, n), M, (S, n), _flags=_lib.FLAG_NO_TMA, **kw),      # cp.async-staged adjoint
    "no_walk": sb.XRayTransform3D((S, n, n), M, (S, n), _flags=_lib.FLAG_NO_WALK, **kw),    # first-generation plane kernels
}
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((S, n, n), device=dev, generator=g)
y = torch.randn((V, S, n), device=dev, generator=g)
upd = float(S) * n * n * V
res = {}
for name, A in ops.items():
    info = A.plan_info()
    for fn, arg, tag in ((A.__call__, x, "fwd"), (A.adj, y, "adj")):
        for _ in range(2):
            out = fn(arg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            out = fn(arg)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        res[(name, tag)] = out
        kern = _lib.KERNEL_NAMES[info["adj_kernel" if tag == "adj" else "fwd_kernel"]]
        print(f"{name:8s} {tag} [{kern:7s}] {ms:9.3f} ms  {upd / ms / 1e6:8.1f} G updates/s  "
              f"(x{n * n * n / (S * n * n):.0f} -> {ms * n / S * 1024 / V:8.1f} ms at {n}^3 x 1024 views)", flush=True)
for tag in ("fwd", "adj"):
    for other in ("no_joint", "no_tma", "no_walk"):
        a, b = res[("default", tag)], res[(other, tag)]
        rel = (torch.linalg.vector_norm((a - b).double()) / torch.linalg.vector_norm(b.double())).item()
        print(f"{tag}: default vs {other} rel-L2 {rel:.2e}")
