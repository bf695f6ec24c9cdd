"""A/B of the 3D separable forward kernels on a z-slab of the headline geometry (1024^2 plane, 1024 views):
CTA-shared-tile joint forward (walk_forward_tile_kernel) against the register-stationary joint forward
(XCT_FLAG_NO_TILE).  Dense input, CUDA events, 3 warm-ups, 5 repetitions; prints ms per application scaled to
the full 1024 slices and the relative difference of the two results."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scico_b200 as sb
from scico_b200 import _lib

n, V, S = 1024, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = sb.matrices_from_euler_angles((n,) * 3, (n, n), "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
z0 = (n - S) // 2
kw = dict(slice_offset=z0, det_row_offset=z0, det_rows_total=n)
x = torch.randn((S, n, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
res = {}
for name, flags in (("tile", 0), ("registers", _lib.FLAG_NO_TILE)):
    A = sb.XRayTransform3D((S, n, n), M, (S, n), _flags=flags, **kw)
    for _ in range(3):
        y = A(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = A(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[name] = y
    print(f"{name:10s} {ms:8.3f} ms per {S}-slice slab = {ms * n / S:7.1f} ms per 1024^3 application", flush=True)
d = (torch.linalg.vector_norm((res["tile"] - res["registers"]).double()) / torch.linalg.vector_norm(res["registers"].double())).item()
print("tile vs registers rel-L2:", d)
