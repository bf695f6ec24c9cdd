"""The reference's 3D CT example (examples/scripts/ct_3d_tv_padmm.py:42-120) rebuilt for the notebook-pin tests:
tangle phantom (scico/examples.py:529-581, restated), geometry through the ASTRA-free converters, the example's
parameters, and the statistics its ProximalADMM prints (_padmm.py:148-177, 294-345 with fast_dual_residual=True)."""
import numpy as np

NX, NY, NZ, NVIEWS = 128, 256, 64, 10
ALPHA, LAM, RHO, MAXITER = 1e2, 2e0, 5e-3, 1000


def tangle_phantom(nx=NX, ny=NY, nz=NZ):
    xs = 1.0 * np.linspace(-1.0, 1.0, nx, dtype=np.float32)
    ys = 1.0 * np.linspace(-1.0, 1.0, ny, dtype=np.float32)
    zs = 1.0 * np.linspace(-1.0, 1.0, nz, dtype=np.float32)
    xx, yy, zz = np.meshgrid(ys, zs, xs, copy=True)  # (nz, ny, nx)
    xx, yy, zz = 3.0 * xx, 3.0 * yy, 3.0 * zz
    v = (xx * xx * xx * xx - 5.0 * xx * xx + yy * yy * yy * yy - 5.0 * yy * yy + zz * zz * zz * zz - 5.0 * zz * zz
         + 11.8) * 0.2 + 0.5
    v[v <= 2.0] = 2.0 - v[v <= 2.0]
    v[v > 2.0] = 0.0
    v[v < 0.0] = 0.0
    return v


def geometry():
    import scico_b200 as sb

    angles = np.linspace(0, np.pi, NVIEWS, endpoint=False)
    det_count = (NZ, max(NX, NY))
    vectors = sb.angle_to_vector([1.0, 1.0], angles)
    matrices = sb.convert_to_scico_geometry(input_shape=(NZ, NY, NX), det_count=det_count, vectors=vectors)
    return (NZ, NY, NX), matrices, det_count


def snr_db(ref, cmp):
    ref, cmp = np.asarray(ref, np.float64), np.asarray(cmp, np.float64)
    return 10.0 * np.log10(np.var(ref) / np.mean((ref - cmp) ** 2))  # scico/metric.py:34-62


def mae(ref, cmp):
    return float(np.mean(np.abs(np.asarray(ref, np.float64) - np.asarray(cmp, np.float64))))
