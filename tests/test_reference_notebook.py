"""The ORACLE against the reference's own published run (real JAX / XLA) of examples/scripts/ct_3d_tv_padmm.py.

`tests/golden/nb_ct_3d_tv_padmm.npz` holds the iteration statistics the reference printed in
`data/notebooks/ct_3d_tv_padmm.ipynb` (see `tests/golden/make_notebook_golden.py`).  The oracle's projector pair
(`oracle/xray_c.c`) and its ProximalADMM restatement (`oracle/tv_np.py`) rebuild the example on the CPU and must print the
same numbers for the first iterations (the CPU port needs ~1 s per projection, so 6 of the 1000 iterations run here;
`tests/test_gpu_reference_notebook.py` runs all 1000 on the CUDA path).  This pins the oracle on output of the real
reference, not only on the reference source executed over the NumPy stand-in.

`mu = 1.01 ||(C; alpha D)||^2` comes from a power iteration with a random start in the reference; the value used here is the
estimate of 100 power iterations on the CUDA path (another 100-iteration estimate from the oracle: 118984).  Iteration 0
does not depend on it."""
import os

import numpy as np

import _ct3d_example as E
from oracle import tv_np as T
from oracle import xray_c as C

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nb_ct_3d_tv_padmm.npz")
NORM_SQ = 119042.16


def test_oracle_reproduces_the_first_iterations_of_the_reference_notebook():
    g = np.load(GOLD)
    N, M, D = E.geometry()
    M32 = np.asarray(M, np.float32)
    Ao = lambda x: C.project_3d(x, M32, D)  # noqa: E731
    ATo = lambda y: C.back_project_3d(y, M32, N)  # noqa: E731
    y = Ao(E.tangle_phantom())
    mu, nu = 1.01 * NORM_SQ, 1.01
    x, z, u, uo = T.padmm_tv_init(N, y.shape)
    f64 = np.float64
    for it in range(6):
        zo = (z[0].copy(), z[1].copy())
        x, z, u, uo = T.padmm_tv_step(x, z, u, uo, Ao, ATo, y, E.LAM, E.ALPHA, E.RHO, mu, nu)
        cx = (Ao(x), np.float32(E.ALPHA) * T.finite_difference(x))
        pr = np.sqrt(np.sum((cx[0].astype(f64) - z[0]) ** 2) + np.sum((cx[1].astype(f64) - z[1]) ** 2))  # ||A x + B z||
        du = np.sqrt(np.sum((z[0].astype(f64) - zo[0]) ** 2) + np.sum((z[1].astype(f64) - zo[1]) ** 2))  # fast dual residual
        obj = 0.5 * np.sum((z[0].astype(f64) - y) ** 2) + (E.LAM / E.ALPHA) * np.sum(np.sqrt(np.sum(z[1].astype(f64) ** 2, 0)))
        tol = 5e-4 if it == 0 else 2e-3
        for got, key in ((obj, "objective"), (pr, "prml_rsdl"), (du, "dual_rsdl")):
            assert abs(got - g[key][it]) <= tol * g[key][it], (it, key, got, g[key][it])
