"""Per-config timings for BASELINE.md section 4 (run under gpurun on one B200).

Times forward and adjoint of BASELINE.json configs C2 (2D 512^2 x 360), C3 (2D 4096^2 x 2048),
C4 (3D 512^3 x 720, det 512^2) and a tilted-geometry 3D case (general kernels), with CUDA events,
3 warm-ups, and reports voxel-view updates/s plus the parity numbers against the oracle where the
oracle finishes in seconds."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root
import numpy as np
import torch

import scico_b200 as sb
from oracle import xray_c as C
from oracle import xray_np as O

dev = "cuda:0"


def timeit(fn, reps):
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


def flush_l2():
    torch.empty(64 * 1024 * 1024, device=dev).fill_(0.0)  # 256 MB > 126 MB L2


def run(name, A, updates, reps, small_working_set=False):
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(A.input_shape, device=dev, generator=g)
    y = torch.randn(A.output_shape, device=dev, generator=g)
    if small_working_set:  # working set < L2: flush between iterations, time one call at a time
        def one(fn, arg):
            ts = []
            for _ in range(reps + 3):
                flush_l2()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(arg); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return float(np.mean(ts[3:]))
        f_ms, a_ms = one(A, x), one(A.adj, y)
    else:
        f_ms, a_ms = timeit(lambda: A(x), reps), timeit(lambda: A.adj(y), reps)
    Ax, ATy = A(x), A.adj(y)
    a = torch.sum(Ax.double() * y.double()).item()
    b = torch.sum(x.double() * ATy.double()).item()
    gap = abs(a - b) / (torch.linalg.vector_norm(Ax.double()).item() * torch.linalg.vector_norm(y.double()).item())
    info = A.plan_info()
    rec = {"config": name, "path": info["path_name"], "fwd_kernel": info["fwd_kernel"], "adj_kernel": info["adj_kernel"],
           "fwd_ms": f_ms, "adj_ms": a_ms, "fwd_updates_per_s": updates / f_ms * 1e3, "adj_updates_per_s": updates / a_ms * 1e3,
           "pair_updates_per_s": 2 * updates / (f_ms + a_ms) * 1e3, "adjoint_gap": gap,
           "l2": "flushed between calls" if small_working_set else "working set > L2"}
    return rec, x, y, Ax, ATy


out = []
# C2
ang = np.linspace(0, np.pi, 360, endpoint=False)
A = sb.XRayTransform2D((512, 512), ang)
rec, x, y, Ax, ATy = run("C2 2D 512^2 x 360 views, 725 bins", A, 512 * 512 * 360, 10, small_working_set=True)
T = A.view_table
rec["rel_l2_fwd"] = O.rel_l2(Ax.cpu().numpy(), C.project_2d(x.cpu().numpy(), T, A.ny))
rec["rel_l2_adj"] = O.rel_l2(ATy.cpu().numpy(), C.back_project_2d(y.cpu().numpy(), T, (512, 512)))
out.append(rec); print(json.dumps(rec), flush=True)
# C3
ang = np.linspace(0, np.pi, 2048, endpoint=False)
A = sb.XRayTransform2D((4096, 4096), ang)
rec, *_ = run("C3 2D 4096^2 x 2048 views, 5793 bins (1 GPU)", A, 4096 * 4096 * 2048, 3)
out.append(rec); print(json.dumps(rec), flush=True)
# C4
n, V = 512, 720
M = sb.matrices_from_euler_angles((n,) * 3, (n, n), "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
A = sb.XRayTransform3D((n,) * 3, M, (n, n))
rec, *_ = run("C4 3D 512^3 x 720 views, det 512^2 (1 GPU)", A, n ** 3 * V, 3)
out.append(rec); print(json.dumps(rec), flush=True)
# tilted geometry -> general kernels
n, V = 256, 64
angs = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, np.deg2rad(74.0))], 1)
M = sb.matrices_from_euler_angles((n,) * 3, (n + 64, n + 64), "XY", angs)
A = sb.XRayTransform3D((n,) * 3, M, (n + 64, n + 64))
rec, *_ = run("3D 256^3 x 64 views, XY tilt 74 deg (general kernels)", A, n ** 3 * V, 3)
out.append(rec); print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs_r01.json", "w"), indent=1)
