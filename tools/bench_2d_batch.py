"""Batched 2D projector pair (the reference's vmap use: scico/flax/examples/data_generation.py:153-186, flax/inverse.py):
a batch through one call against the same images one by one.  CUDA events, L2-sized or larger working sets.
Usage: python tools/bench_2d_batch.py [n] [views] [batch]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch

    import scico_b200 as sb

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    V = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    dev = "cuda:0"
    A = sb.XRayTransform2D((n, n), np.linspace(0, np.pi, V, endpoint=False))
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((B, n, n), device=dev, generator=g)
    y = torch.randn((B,) + A.output_shape, device=dev, generator=g)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    out_y, out_x = torch.empty_like(y), torch.empty_like(x)
    res = {"shape": f"{B} images of {n}^2, {V} views, {A.ny} bins", "updates_per_application": float(B) * n * n * V}
    res["fwd_ms_batched"] = timed(lambda: A.project(x, out=out_y))
    res["adj_ms_batched"] = timed(lambda: A.back_project(y, out=out_x))
    xb = out_x.clone()
    yb = out_y.clone()
    res["fwd_ms_one_by_one"] = timed(lambda: [A.project(x[i], out=out_y[i]) for i in range(B)])
    res["adj_ms_one_by_one"] = timed(lambda: [A.back_project(y[i], out=out_x[i]) for i in range(B)])
    res["adj_batched_equals_one_by_one_bitwise"] = bool(torch.equal(xb, out_x))
    res["adj_batched_vs_one_by_one_rel_l2"] = float(torch.linalg.vector_norm(xb - out_x) / torch.linalg.vector_norm(out_x))
    res["adj_batched_vs_one_by_one_max_abs"] = float((xb - out_x).abs().max())
    res["adj_fraction_of_pixels_that_differ"] = float((xb != out_x).float().mean())
    res["fwd_batched_vs_one_by_one_rel_l2"] = float(torch.linalg.vector_norm(yb - out_y) / torch.linalg.vector_norm(out_y))
    u = res["updates_per_application"]
    for k in ("fwd_ms_batched", "adj_ms_batched", "fwd_ms_one_by_one", "adj_ms_one_by_one"):
        res[k.replace("_ms_", "_updates_per_s_")] = u / (res[k] * 1e-3)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
