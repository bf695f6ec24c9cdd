"""Solver timings for BASELINE.json's second metric ("TV-ADMM CT iters/s"), run under gpurun on one B200.

C1 (the reference example shape, examples/scripts/ct_tv_admm.py scaled to 256^2 x 180 views): TV-regularised
ADMM with a CG x-step, lam = 2, rho = 5, 25 iterations, CG tol 1e-4 / maxiter 25, x0 = clip(fbp(y), 0, 1).
Timed on the GPU (TVADMM: CUDA projector pair + fused kernels, wall clock including the per-CG-iteration
scalar read-back) and, beside it, the oracle's ADMM (NumPy vector arithmetic + the C port of the projector on all
host cores) on the same problem; the reconstructions are compared.
3D (C4 operator, 512^3 x 720 views): PDHG / proximal ADMM / linearised ADMM iterations per second.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root
import numpy as np
import torch

import scico_b200 as sb
from scico_b200.optimize import TVADMM, TVPDHG, TVLinearizedADMM, TVProximalADMM
from oracle import tv_np as T
from oracle import xray_c as C
from oracle import xray_np as O

dev = "cuda:0"
out = []


def disc_phantom(n, seed=1234, discs=40):
    """Seeded union of discs (stand-in for xdesign's Foam, which is not installable here)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float32) / n - 0.5
    img = ((xx ** 2 + yy ** 2) < 0.45 ** 2).astype(np.float32)
    for _ in range(discs):
        cx, cy = rng.uniform(-0.3, 0.3, 2)
        r = rng.uniform(0.01, 0.06)
        img[(xx - cx) ** 2 + (yy - cy) ** 2 < r * r] = 0.0
    return img


# ---- C1: 2D TV-ADMM ---------------------------------------------------------------------------
n, V = 256, 180
angles = np.linspace(0, np.pi, V, endpoint=False)
A = sb.XRayTransform2D((n, n), angles)
x_gt = disc_phantom(n)
xt = torch.as_tensor(x_gt, device=dev)
y = A(xt)
lam, rho, iters, cg_tol, cg_max = 2.0, 5.0, 25, 1e-4, 25
x0 = torch.clamp(A.fbp(y), 0.0, 1.0)


def run_gpu():
    S = TVADMM(A, y, lam, rho, x0=x0, maxiter=iters, cg_tol=cg_tol, cg_maxiter=cg_max)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    S.solve()
    torch.cuda.synchronize()
    return S, time.perf_counter() - t0


run_gpu()  # warm-up (plan creation, allocator)
S, sec = run_gpu()
snr = lambda ref, rec: 10 * np.log10(np.sum(ref ** 2) / np.sum((ref - rec) ** 2))  # noqa: E731
rec_gpu = np.clip(S.x.cpu().numpy(), 0, 1)
c1 = {"config": "C1 2D 256^2 x 180 views, TV-ADMM (lam 2, rho 5, 25 iters, CG tol 1e-4 / maxiter 25, x0 = clip(fbp))",
      "gpu_seconds": sec, "gpu_admm_iters_per_s": iters / sec, "gpu_cg_iters_total": S.cg_iters_total,
      "gpu_projector_pairs_per_s": (S.cg_iters_total + iters) / sec,
      "snr_fbp_db": float(snr(x_gt, x0.cpu().numpy())), "snr_tv_db": float(snr(x_gt, rec_gpu))}
# CPU arm: the oracle's ADMM on the same data (C port of the projector, all host threads)
Tb = O.view_table_2d(angles, A.x0, A.dx, A.y0)
Ao = lambda v: C.project_2d(v, Tb, A.ny)  # noqa: E731
ATo = lambda v: C.back_project_2d(v, Tb, (n, n))  # noqa: E731
yn, x0n = y.cpu().numpy(), x0.cpu().numpy()
x, z, u = T.admm_tv_init(x0n)
t0 = time.perf_counter()
cg_total = 0
for _ in range(iters):
    x, z, u, info = T.admm_tv_step(x, z, u, Ao, ATo, yn, lam, rho, cg_tol=cg_tol, cg_maxiter=cg_max)
    cg_total += info["num_iter"]
cpu_sec = time.perf_counter() - t0
c1.update({"cpu_seconds": cpu_sec, "cpu_admm_iters_per_s": iters / cpu_sec, "cpu_cg_iters_total": cg_total,
           "cpu_threads": C.num_threads(), "cpu_kind": "port (oracle/tv_np.py + oracle/xray_c.c)",
           "rel_l2_gpu_vs_cpu": O.rel_l2(S.x.cpu().numpy(), x)})
out.append(c1)
print(json.dumps(c1), flush=True)

# ---- 3D: C4 operator, iterations per second of the device-resident solvers -----------------------
n, V = 512, 720
M = sb.matrices_from_euler_angles((n,) * 3, (n, n), "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
A3 = sb.XRayTransform3D((n,) * 3, M, (n, n))
lin = torch.linspace(-3.0, 3.0, n, device=dev)
zz, yy, xx = lin[:, None, None], lin[None, :, None], lin[None, None, :]
val = (xx ** 4 - 5 * xx ** 2 + yy ** 4 - 5 * yy ** 2 + zz ** 4 - 5 * zz ** 2 + 11.8) * 0.2 + 0.5
vol = torch.where(val <= 2.0, 2.0 - val, torch.zeros_like(val)).clamp_(min=0.0).contiguous()
y3 = A3(vol)
del val


def time_steps(S, k):
    S.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        S.step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


alpha = 1e2
mu_p, nu_p = TVProximalADMM.estimate_parameters(A3, alpha=alpha, maxiter=10)
mu_l, nu_l = TVLinearizedADMM.estimate_parameters(A3, maxiter=10)
tau, sigma = TVPDHG.estimate_parameters(A3, maxiter=10)
for name, mk, k in (
    ("PDHG", lambda: TVPDHG(A3, y3, 0.1, tau, sigma), 5),
    ("ProximalADMM (alpha 1e2, rho 5e-3)", lambda: TVProximalADMM(A3, y3, 2.0, 5e-3, mu_p, nu_p, alpha=alpha), 5),
    ("LinearizedADMM", lambda: TVLinearizedADMM(A3, y3, 0.1, mu_l, nu_l), 5),
    ("ADMM + CG (3 CG iterations per x-step)", lambda: TVADMM(A3, y3, 2.0, 5.0, cg_tol=1e-30, cg_maxiter=3), 2),
):
    Sx = mk()
    ms = time_steps(Sx, k)
    rec = {"config": f"C4 3D 512^3 x 720 views, det 512^2: {name}", "ms_per_iter": ms, "iters_per_s": 1e3 / ms}
    out.append(rec)
    print(json.dumps(rec), flush=True)
    del Sx
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/solvers_r01.json", "w"), indent=1)
