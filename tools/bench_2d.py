"""Timing of the 2D pair at BASELINE.json configs[2] (C3: 4096^2 image, 2048 views, 5793 bins; or `n views` from the
command line), dense input, CUDA events, 3 warm-ups, 5 repetitions.  Used under ncu for the 2D kernels' counters."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scico_b200 as sb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
V = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
A = sb.XRayTransform2D((n, n), np.linspace(0, np.pi, V, endpoint=False))
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(A.input_shape, device="cuda", generator=g)
y = torch.randn(A.output_shape, device="cuda", generator=g)
for name, f, arg in (("forward", A.project, x), ("adjoint", A.back_project, y)):
    for _ in range(3):
        f(arg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f(arg)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name} {ms:8.3f} ms = {n * n * V / ms / 1e9:6.3f}e12 updates/s", flush=True)
