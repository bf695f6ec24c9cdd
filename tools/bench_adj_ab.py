"""Timing of the 3D separable adjoint on a z-slab of the headline geometry (1024^2 plane, 1024 views, dense input,
CUDA events, 3 warm-ups, 5 repetitions), scaled to the full 1024 slices; another build of the library is selected
with SCICO_B200_LIB; the scalar-tap walk adjoint (XCT_FLAG_NO_ADJ_VEC) and the plane adjoint (XCT_FLAG_NO_WALK) of the
same library are timed and compared beside it."""
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
y = torch.randn((V, S, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
A = sb.XRayTransform3D((S, n, n), M, (S, n), **kw)
for _ in range(3):
    x = A.T(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    x = A.T(y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"adjoint {ms:8.3f} ms per {S}-slice slab = {ms * n / S:7.1f} ms per 1024^3 application", flush=True)
print("plan:", {k: A.analyse().get(k) for k in ("adj_kernel", "adj_tma", "adj_interleaved")})
for name, flag in (("scalar taps", getattr(_lib, "FLAG_NO_ADJ_VEC", None)), ("plane", _lib.FLAG_NO_WALK)):
    if flag is None:
        continue
    B = sb.XRayTransform3D((S, n, n), M, (S, n), _flags=flag, **kw)
    for _ in range(2):
        xr = B.T(y)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        xr = B.T(y)
    e1.record()
    torch.cuda.synchronize()
    d = (torch.linalg.vector_norm((x - xr).double()) / torch.linalg.vector_norm(xr.double())).item()
    print(f"{name:12s} {e0.elapsed_time(e1) / 3 * n / S:7.1f} ms per 1024^3 application; rel-L2 to the default: {d}")
