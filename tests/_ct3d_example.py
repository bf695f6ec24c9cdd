"""The reference's 3D CT example (examples/scripts/ct_3d_tv_padmm.py:42-120) rebuilt for the notebook-pin tests:
tangle phantom (scico/examples.py:529-581, restated), geometry through the ASTRA-free converters, the example's
parameters, and the statistics its ProximalADMM prints (_padmm.py:148-177, 294-345 with fast_dual_residual=True)."""
import numpy as np

NX, NY, NZ, NVIEWS = 128, 256, 64, 10
ALPHA, LAM, RHO, MAXITER = 1e2, 2e0, 5e-3, 1000


def tangle_phantom(nx=NX, ny=NY, nz=NZ):
    """(nz, ny, nx) float32 tanglecube phantom: p = 2 - t where 0 <= t <= 2, else 0, with
    t = 0.2 (x^4 - 5 x^2 + y^4 - 5 y^2 + z^4 - 5 z^2 + 11.8) + 0.5 on [-3, 3]^3; float32 throughout and summed in the
    reference's order (first the axis of length ny, then nz, then nx: its meshgrid naming), so the values are the
    reference's bit for bit."""
    def quartic(n):
        t = np.float32(3.0) * np.linspace(-1.0, 1.0, n, dtype=np.float32)
        return t * t * t * t, np.float32(5.0) * t * t

    (qy, sy), (qz, sz), (qx, sx) = quartic(ny), quartic(nz), quartic(nx)
    acc = np.broadcast_to(qy[None, :, None], (nz, ny, nx)) - sy[None, :, None]
    acc = acc + qz[:, None, None]
    acc = acc - sz[:, None, None]
    acc = acc + qx[None, None, :]
    acc = acc - sx[None, None, :]
    t = (acc + np.float32(11.8)) * np.float32(0.2) + np.float32(0.5)
    p = np.float32(2.0) - t
    # the reference flips t <= 2 to 2 - t FIRST and then zeroes everything above 2: t < 0 (p > 2) ends up 0 as well
    return np.where((t <= 2.0) & (p <= 2.0), p, np.float32(0.0)).astype(np.float32)


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
