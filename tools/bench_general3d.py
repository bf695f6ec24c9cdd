"""General-matrix 3D projector (brick kernels, xct_brick.cuh) against the thread-per-voxel kernels they replace:
256^3 volume, 64 views, 74-degree XY tilt (examples/scripts/ct_projector_comparison_3d.py:44-51 scaled up),
detector 320 x 320; plus a 512^3 x 64 case.  CUDA events, 3 warm-ups, 5 repetitions; checks that the two
families agree.  Run under gpurun on one B200; writes gpurun_out/general3d.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scico_b200 as sb
from scico_b200 import _lib

dev = "cuda:0"


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


out = []
CONFIGS = ((256, 64, "XY", "74 deg XY tilt"), (512, 64, "XY", "74 deg XY tilt"), (256, 64, "XYZ", "random orientations"))
FAMILIES = (("brick", 0), ("brick_no_tma", _lib.FLAG_NO_TMA), ("thread_per_voxel", _lib.FLAG_NO_BRICK))
if "--profile" in sys.argv:  # one configuration, brick kernels only (the workload of an ncu capture)
    CONFIGS, FAMILIES = CONFIGS[:1], FAMILIES[:1]
for n, V, seq, label in CONFIGS:
    D = (n + 64, n + 64)
    if seq == "XY":
        ang = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, np.deg2rad(74.0))], 1)
    else:
        ang = np.random.default_rng(5).uniform(0, 2 * np.pi, size=(V, 3))
    M = sb.matrices_from_euler_angles((n,) * 3, D, seq, ang)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn((n,) * 3, device=dev, generator=g)
    y = torch.randn((V,) + D, device=dev, generator=g)
    rec = {"config": f"3D {n}^3 x {V} views, {label}, det {D[0]}x{D[1]}", "updates": float(n) ** 3 * V}
    res = {}
    for name, flags in FAMILIES:
        A = sb.XRayTransform3D((n,) * 3, M, D, _flags=flags)
        info = A.plan_info()
        f_ms = timeit(lambda: A(x))
        a_ms = timeit(lambda: A.adj(y))
        res[name] = (A(x), A.adj(y))
        rec[name] = {"fwd_ms": f_ms, "adj_ms": a_ms, "fwd_updates_per_s": rec["updates"] / f_ms * 1e3,
                     "adj_updates_per_s": rec["updates"] / a_ms * 1e3, "pair_updates_per_s": 2 * rec["updates"] / (f_ms + a_ms) * 1e3,
                     "fwd_kernel": info["fwd_kernel"], "adj_kernel": info["adj_kernel"], "adj_tma": info["adj_tma"],
                     "brick_views": A.analyse()["brick_views"]}
    rel = lambda a, b: (torch.linalg.vector_norm((a - b).double()) / torch.linalg.vector_norm(b.double())).item()  # noqa: E731
    if "thread_per_voxel" in res:
        rec["brick_vs_thread_per_voxel_rel_l2"] = {"fwd": rel(res["brick"][0], res["thread_per_voxel"][0]),
                                                   "adj": rel(res["brick"][1], res["thread_per_voxel"][1])}
    Ax, ATy = res["brick"]
    rec["adjoint_gap"] = abs(torch.sum(Ax.double() * y.double()).item() - torch.sum(x.double() * ATy.double()).item()) / (
        torch.linalg.vector_norm(Ax.double()).item() * torch.linalg.vector_norm(y.double()).item())
    out.append(rec)
    print(json.dumps(rec), flush=True)
    del res, x, y
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/general3d.json", "w"), indent=1)
