"""C2 (BASELINE.json configs[1]: 2D 512^2, 360 views, 725 bins) forward / adjoint, L2 flushed before every call:
one launch for all view classes (default) against one launch per class (XCT_FLAG_2D_PER_CLASS)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scico_b200 as sb
from scico_b200 import _lib

dev = "cuda:0"
n, V = 512, 360
ang = np.linspace(0, np.pi, V, endpoint=False)
x = torch.randn((n, n), device=dev)
for name, flags in (("one launch", 0), ("per class", _lib.FLAG_2D_PER_CLASS)):
    A = sb.XRayTransform2D((n, n), ang, _flags=flags)
    y = A(x)
    res = {}
    for tag, fn, arg in (("fwd", A.__call__, x), ("adj", A.adj, y)):
        for _ in range(3):
            fn(arg)
        ts = []
        for _ in range(20):
            torch.empty(64 * 1024 * 1024, device=dev).fill_(0.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(arg); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[tag] = float(np.median(ts))
    upd = n * n * V
    print(f"{name:11s} fwd {res['fwd']*1e3:7.1f} us  adj {res['adj']*1e3:7.1f} us  pair {2*upd/(res['fwd']+res['adj'])*1e3:.3e} updates/s", flush=True)
