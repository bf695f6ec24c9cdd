"""Developer GPU check: parity vs the oracle on a few shapes + first timings (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root
import numpy as np, torch
import scico_b200 as sb
from scico_b200 import _lib
from scico_b200.xray import debug_weights_3d, debug_weights_2d
from oracle import xray_np as O, xray_c as C

rng = np.random.default_rng(0)
dev = "cuda:0"

def t2n(t): return t.cpu().numpy()

def check3d(name, N, D, M, flags=0, big=False):
    A = sb.XRayTransform3D(N, M, D, _flags=flags)
    info = A.plan_info()
    x = rng.standard_normal(N).astype(np.float32); y = rng.standard_normal(A.output_shape).astype(np.float32)
    Ax = t2n(A(torch.as_tensor(x, device=dev))); ATy = t2n(A.adj(torch.as_tensor(y, device=dev)))
    rAx = C.project_3d(x, A.matrices, D, fused=big); rATy = C.back_project_3d(y, A.matrices, N)
    ns, ref = O.adjoint_gap(Ax, y, x, ATy)
    wok = ""
    if not big:
        for v in range(min(3, len(M))):
            ul, w = debug_weights_3d(A, v)
            rul, rw = C.weights_3d(A.matrices[v], N, D)
            wok += f" v{v}:ul={np.array_equal(ul, rul)},zeros={np.array_equal(w==0, rw==0)},dw={np.abs(w-rw).max():.1e}"
    print(f"[3D {name}] path={info['path_name']} gs={info['fwd_lane_stride']} fwd rel-L2 {O.rel_l2(Ax, rAx):.2e} adj rel-L2 {O.rel_l2(ATy, rATy):.2e} adjgap {ns:.1e}/{ref:.1e}{wok}", flush=True)

def check2d(name, nx, angles, flags=0, **kw):
    A = sb.XRayTransform2D(nx, angles, _flags=flags, **kw)
    info = A.plan_info()
    x = rng.standard_normal(nx).astype(np.float32); y = rng.standard_normal(A.output_shape).astype(np.float32)
    Ax = t2n(A(torch.as_tensor(x, device=dev))); ATy = t2n(A.adj(torch.as_tensor(y, device=dev)))
    T = O.view_table_2d(angles, A.x0, A.dx, A.y0)
    assert np.array_equal(T, A.view_table)
    rAx = C.project_2d(x, T, A.ny); rATy = C.back_project_2d(y, T, nx)
    ns, ref = O.adjoint_gap(Ax, y, x, ATy)
    inds, w = debug_weights_2d(A, len(angles)//3); rinds, rw = C.weights_2d(T[len(angles)//3], nx)
    print(f"[2D {name}] path={info['path_name']} gs={info['fwd_lane_stride']} fwd rel-L2 {O.rel_l2(Ax, rAx):.2e} adj rel-L2 {O.rel_l2(ATy, rATy):.2e} adjgap {ns:.1e}/{ref:.1e} inds={np.array_equal(inds, rinds)} dw={np.abs(w-rw).max():.1e}", flush=True)

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3

print(torch.cuda.get_device_name(0), flush=True)
# ---- 3D parity
for N, D, V in [((16,16,16),(11,11),3), ((17,18,19),(20,21),5), ((40,70,50),(40,90),7), ((33,130,129),(33,190),9)]:
    ang = np.linspace(0, np.pi, V, endpoint=False)[:, None]
    M = sb.matrices_from_euler_angles(N, D, "X", ang)
    check3d(f"X {N}", N, D, M)
    check3d(f"X {N} forced-general", N, D, M, flags=_lib.FLAG_FORCE_GENERAL)
N, D = (17,18,19), (20,21)
ang = np.stack([np.linspace(0,np.pi,5,endpoint=False), np.full(5,np.deg2rad(74))],1)
check3d("XY tilt", N, D, sb.matrices_from_euler_angles(N, D, "XY", ang))
M = sb.matrices_from_euler_angles((32,48,40),(32,70),"X",np.linspace(0,np.pi,6,endpoint=False)[:,None]); M[:,1,3] += 0.25; M[:,0,3] += 0.25
check3d("X quirk offsets", (32,48,40), (32,70), M)
M = sb.matrices_from_euler_angles((32,48,40),(40,70),"X",np.linspace(0,np.pi,6,endpoint=False)[:,None], voxel_spacing=[0.8,1,1])
check3d("X unaligned rows", (32,48,40), (40,70), M)
# mid-size
N=(64,256,256); D=(64,300); V=32
M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0,np.pi,V,endpoint=False)[:,None])
check3d("X 64x256x256", N, D, M, big=True)
# ---- 2D parity
check2d("12x13", (12,13), np.linspace(0,np.pi,10,endpoint=False))
check2d("16x16 det11", (16,16), np.linspace(0,np.pi,3,endpoint=False), det_count=11, dx=1/np.sqrt(2))
check2d("64x64", (64,64), np.linspace(0,np.pi,90,endpoint=False))
check2d("100x130 general", (100,130), np.linspace(0,np.pi,45,endpoint=False), flags=_lib.FLAG_FORCE_GENERAL)
check2d("512x512", (512,512), np.linspace(0,np.pi,360,endpoint=False))
check2d("300x200 dx=.5", (300,200), np.linspace(0,2*np.pi,50,endpoint=False), dx=0.5)
# ---- timings
for N0, N, V in [(64, 512, 720), (128, 1024, 256)]:
    Ns=(N0,N,N); D=(N0,N)
    M = sb.matrices_from_euler_angles(Ns, D, "X", np.linspace(0,np.pi,V,endpoint=False)[:,None])
    A = sb.XRayTransform3D(Ns, M, D)
    x = torch.rand(Ns, device=dev); y = torch.rand(A.output_shape, device=dev)
    upd = np.prod(Ns) * V
    tf = timeit(lambda: A(x)); ta = timeit(lambda: A.adj(y))
    print(f"[time 3D sep {Ns} V={V}] fwd {tf*1e3:.2f} ms {upd/tf:.3e} upd/s | adj {ta*1e3:.2f} ms {upd/ta:.3e} upd/s | model-HBM frac fwd {4*upd/tf/6548.5e9:.2f} adj {4*upd/ta/6548.5e9:.2f}", flush=True)
for n, V in [(512, 360), (4096, 256)]:
    A = sb.XRayTransform2D((n,n), np.linspace(0,np.pi,V,endpoint=False))
    x = torch.rand((n,n), device=dev); y = torch.rand(A.output_shape, device=dev)
    upd = n*n*V
    tf = timeit(lambda: A(x), 5); ta = timeit(lambda: A.adj(y), 5)
    print(f"[time 2D {n} V={V}] fwd {tf*1e3:.3f} ms {upd/tf:.3e} upd/s | adj {ta*1e3:.3f} ms {upd/ta:.3e} upd/s", flush=True)
Ns=(64,128,128); D=(64,128); V=64
M = sb.matrices_from_euler_angles(Ns, D, "X", np.linspace(0,np.pi,V,endpoint=False)[:,None])
A = sb.XRayTransform3D(Ns, M, D, _flags=_lib.FLAG_FORCE_GENERAL)
x = torch.rand(Ns, device=dev); y = torch.rand(A.output_shape, device=dev); upd = np.prod(Ns)*V
tf = timeit(lambda: A(x)); ta = timeit(lambda: A.adj(y))
print(f"[time 3D general {Ns} V={V}] fwd {upd/tf:.3e} upd/s | adj {upd/ta:.3e} upd/s", flush=True)
