import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import scico_b200 as sb
from scico_b200 import _lib
which = sys.argv[1]
N, D = (21, 30, 37), (44, 48)
ang = np.stack([np.linspace(0, np.pi, 7, endpoint=False), np.full(7, np.deg2rad(74.0))], 1)
M = sb.matrices_from_euler_angles(N, D, "XY", ang)
A = sb.XRayTransform3D(N, M, D)
print(A.plan_info(), A.analyse()["brick_views"], flush=True)
x = torch.randn(N, device="cuda"); y = torch.randn(A.output_shape, device="cuda")
if which == "fwd":
    r = A(x); torch.cuda.synchronize(); print("fwd ok", float(r.abs().sum()))
else:
    r = A.adj(y); torch.cuda.synchronize(); print("adj ok", float(r.abs().sum()))
