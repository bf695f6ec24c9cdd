"""The ORACLE (CPU: C projector pair + oracle/tv_np.py) on the reference's ct_3d_tv_padmm example for many iterations,
against the statistics table of the reference's executed notebook (tests/golden/nb_ct_3d_tv_padmm.npz).  The CPU suite
runs 6 iterations of this (tests/test_reference_notebook.py); this script runs 120 in about 3.5 minutes.  Relative
deviation from the printed numbers per statistic; usage: python tools/oracle_notebook_long.py [iterations]"""
import json
import os
import sys
import time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import _ct3d_example as E
from oracle import tv_np as T, xray_c as C
C.build()
N, M, D = E.geometry()
M32 = np.asarray(M, np.float32)
Ao = lambda x: C.project_3d(x, M32, D)
ATo = lambda y: C.back_project_3d(y, M32, N)
y = Ao(E.tangle_phantom())
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'nb_ct_3d_tv_padmm.npz'))
mu, nu = 1.01*119042.16, 1.01
x, z, u, uo = T.padmm_tv_init(N, y.shape)
f64=np.float64
dev={'objective':[], 'prml_rsdl':[], 'dual_rsdl':[]}
t0=time.time()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 120):
    zo=(z[0].copy(), z[1].copy())
    x, z, u, uo = T.padmm_tv_step(x, z, u, uo, Ao, ATo, y, E.LAM, E.ALPHA, E.RHO, mu, nu)
    cx=(Ao(x), np.float32(E.ALPHA)*T.finite_difference(x))
    pr=np.sqrt(np.sum((cx[0].astype(f64)-z[0])**2)+np.sum((cx[1].astype(f64)-z[1])**2))
    du=np.sqrt(np.sum((z[0].astype(f64)-zo[0])**2)+np.sum((z[1].astype(f64)-zo[1])**2))
    obj=0.5*np.sum((z[0].astype(f64)-y)**2)+(E.LAM/E.ALPHA)*np.sum(np.sqrt(np.sum(z[1].astype(f64)**2,0)))
    for k,v in (('objective',obj),('prml_rsdl',pr),('dual_rsdl',du)):
        dev[k].append(abs(v-g[k][it])/g[k][it])
    if it%20==19: print(it, {k: max(v) for k,v in dev.items()}, time.time()-t0, flush=True)
print(json.dumps({k: {'max': float(max(v)), 'argmax': int(np.argmax(v))} for k, v in dev.items()}))
