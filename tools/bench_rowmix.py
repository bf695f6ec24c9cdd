"""Forward timing for a geometry whose detector rows are finer than the slices (two-row mixing):
joint-column kernel with the ROWS_MIX flush against the one-column walk and the plane kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scico_b200 as sb
from scico_b200 import _lib

n, V, d0 = 512, 360, 683
M = sb.matrices_from_euler_angles((n,) * 3, (d0, n), "X", np.linspace(0, np.pi, V, endpoint=False)[:, None], det_spacing=[0.75, 1.0])
x = torch.rand((n,) * 3, device="cuda")
ref = None
for name, fl in (("joint ROWS_MIX", 0), ("one-column walk", _lib.FLAG_NO_JOINT), ("plane", _lib.FLAG_NO_WALK)):
    A = sb.XRayTransform3D((n,) * 3, M, (d0, n), _flags=fl)
    for _ in range(2):
        y = A(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = A(x)
    y = A(x)
    e1.record()
    torch.cuda.synchronize()
    info = A.plan_info()
    ref = y if ref is None else ref
    rel = (torch.linalg.vector_norm(y - ref) / torch.linalg.vector_norm(ref)).item()
    print(f"{name:16s} fwd_kernel {info['fwd_kernel']} joint {info['fwd_joint']}  {e0.elapsed_time(e1) / 2:8.2f} ms  rel-L2 vs first {rel:.1e}", flush=True)
