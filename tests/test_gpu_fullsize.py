"""Full-size checks on the B200 through size-independent properties (BASELINE.json configs).

The oracle cannot produce a 1024^3 x 1024-view result in seconds, so at full size the CUDA path is
checked by (a) the adjoint identity on white noise, (b) the oracle evaluated at a random SAMPLE of
voxels (back projection) and on a few complete views (forward), (c) linearity, and (d) agreement of
the two independent kernel families (separable plane kernels vs general thread-per-voxel kernels).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scico_b200 as sb
from scico_b200 import _lib
from oracle import xray_c as C
from oracle import xray_np as O

TOL = 1e-5


def _x_mats(N, D, V):
    return sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])


def _dots(torch, Ax, y, x, ATy):
    a = torch.sum(Ax.double() * y.double()).item()
    b = torch.sum(x.double() * ATy.double()).item()
    ns = abs(a - b) / (torch.linalg.vector_norm(Ax.double()).item() * torch.linalg.vector_norm(y.double()).item())
    return ns, abs(a - b) / max(abs(a), abs(b))


def test_c5_1024cubed_1024views_adjoint_identity_and_sampled_oracle(cuda_device):
    """BASELINE.json configs[4] operator: 1024^3 volume, 1024 views, 1024^2 detector."""
    import torch

    n, V = 1024, 1024
    N, D = (n, n, n), (n, n)
    M = _x_mats(N, D, V)
    A = sb.XRayTransform3D(N, M, D)
    assert A.plan_info()["path_name"] == "3d_sep"
    g = torch.Generator(device=cuda_device).manual_seed(0)
    x = torch.randn(N, device=cuda_device, generator=g)
    y = torch.randn(A.output_shape, device=cuda_device, generator=g)
    Ax = A(x)
    ATy = A.adj(y)
    ns, ref = _dots(torch, Ax, y, x, ATy)
    assert ns < TOL and ref < 1e-4, (ns, ref)
    # back projection against the oracle at 4096 random voxels (plus the 8 corners)
    rng = np.random.default_rng(1)
    pts = rng.integers(0, n, size=(4096, 3)).astype(np.int32)
    corners = np.array([[a, b, c] for a in (0, n - 1) for b in (0, n - 1) for c in (0, n - 1)], np.int32)
    pts = np.concatenate([pts, corners])
    want = C.back_project_3d_points(y.cpu().numpy(), A.matrices, pts)
    got = ATy[pts[:, 0], pts[:, 1], pts[:, 2]].cpu().numpy()
    assert O.rel_l2(got, want) <= TOL
    # forward against the oracle on complete views of 4-slice slabs of the same volume (detector rows depend on
    # the slice only, so a slab's rows are complete): the median view of EVERY class the joint forward kernel
    # launches separately (major axis x sign of the minor coefficient x sign of the major coefficient,
    # xct_api.cu analyse_views) and every view that takes the kernel's per-view E2 variant (ViewRec::fjump:
    # major coefficient within rounding distance of 1 -- the views next to theta = 0 and pi/2, both signs)
    ca, cb = A.matrices[:, 1, 1], A.matrices[:, 1, 2]
    major_b = np.abs(cb) >= np.abs(ca)
    mj, mn = np.where(major_b, cb, ca), np.where(major_b, ca, cb)
    umax = np.abs(ca) * n + np.abs(cb) * n + np.abs(A.matrices[:, 1, 3]) + 2.0
    ulp = np.ldexp(1.0, np.floor(np.log2(umax)).astype(int) - 23)
    fjump = np.abs(mj) + 5 * ulp > 1.0
    cls = major_b * 4 + (mn >= 0) * 2 + (mj > 0)
    vs = sorted({int(np.flatnonzero((cls == c) & ~fjump)[np.count_nonzero((cls == c) & ~fjump) // 2])
                 for c in range(8) if np.any((cls == c) & ~fjump)} | set(np.flatnonzero(fjump).tolist()) | {0, V - 1})
    info = A.analyse()
    assert sum(1 for k in info["joint_views"] if k) >= 4 and len(set(cls[~fjump])) >= 4 and 2 <= fjump.sum() <= 64
    assert set(cls[vs]) == set(cls) and {bool(m > 0) for m in mj[fjump]} == {True, False}
    for z0 in (0, 510, n - 4):  # both volume edges and the centre
        z1 = z0 + 4
        want_f = C.project_3d(x[z0:z1].cpu().numpy(), A.matrices[vs], D, slice_offset=z0, fused=True)[:, z0:z1]
        got_f = Ax[vs][:, z0:z1].cpu().numpy()
        per_view = [O.rel_l2(got_f[i], want_f[i]) for i in range(len(vs))]
        assert max(per_view) <= TOL, (z0, vs[int(np.argmax(per_view))], max(per_view))


def test_c4_512cubed_720views_plane_vs_general_and_linearity(cuda_device):
    """BASELINE.json configs[3]: 512^3, 720 views, 512^2 detector (smaller than the diagonal, so
    out-of-bounds masking is exercised).  A 64-slice slab keeps the general kernels' time bounded."""
    import torch

    n, V, S = 512, 720, 64
    N, D = (S, n, n), (S, n)
    M = _x_mats((n, n, n), (n, n), V)
    kw = dict(slice_offset=224, det_row_offset=224, det_rows_total=n)
    A = sb.XRayTransform3D(N, M, D, **kw)
    G = sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_FORCE_GENERAL, **kw)
    assert A.plan_info()["path_name"] == "3d_sep" and G.plan_info()["path_name"] == "3d_general"
    g = torch.Generator(device=cuda_device).manual_seed(2)
    x = torch.randn(N, device=cuda_device, generator=g)
    y = torch.randn(A.output_shape, device=cuda_device, generator=g)
    Ax, ATy = A(x), A.adj(y)
    rel = lambda a, b: (torch.linalg.vector_norm((a - b).double()) / torch.linalg.vector_norm(b.double())).item()
    assert rel(Ax, G(x)) <= TOL
    assert rel(ATy, G.adj(y)) <= TOL
    ns, ref = _dots(torch, Ax, y, x, ATy)
    assert ns < TOL
    x2 = torch.randn(N, device=cuda_device, generator=g)
    assert rel(A(1.5 * x - 0.5 * x2), 1.5 * Ax - 0.5 * A(x2)) <= TOL


def test_c3_4096sq_2048views_2d(cuda_device):
    """BASELINE.json configs[2] operator on one GPU: 4096^2, 2048 views, 5793 bins."""
    import torch

    n, V = 4096, 2048
    angles = np.linspace(0, np.pi, V, endpoint=False)
    A = sb.XRayTransform2D((n, n), angles)
    assert A.ny == 5793 and A.plan_info()["path_name"] == "2d_plane"
    g = torch.Generator(device=cuda_device).manual_seed(3)
    x = torch.randn((n, n), device=cuda_device, generator=g)
    y = torch.randn(A.output_shape, device=cuda_device, generator=g)
    Ax, ATy = A(x), A.adj(y)
    ns, ref = _dots(torch, Ax, y, x, ATy)
    assert ns < TOL, (ns, ref)
    # 6 complete views against the oracle (forward), a view-block sub-operator for the adjoint
    vs = [0, 1, 700, 1024, 1500, 2047]
    T = A.view_table
    want = C.project_2d(x.cpu().numpy(), T[vs], A.ny, fused=True)
    assert O.rel_l2(Ax[vs].cpu().numpy(), want) <= TOL
    B = sb.XRayTransform2D((n, n), angles[1000:1016], det_count=A.ny)
    np.testing.assert_array_equal(B.view_table, T[1000:1016])
    want_a = C.back_project_2d(y[1000:1016].cpu().numpy(), T[1000:1016], (n, n))
    assert O.rel_l2(B.adj(y[1000:1016].contiguous()).cpu().numpy(), want_a) <= TOL
    # mass conservation: default det_count covers the diagonal, nothing falls off
    ones = A(torch.ones((n, n), device=cuda_device))
    assert torch.allclose(ones.sum(dim=1), torch.full((V,), float(n * n), device=cuda_device), rtol=1e-5)


def test_more_than_2_to_the_31_elements_per_array(cuda_device):
    """Volume and sinogram of 2112 x 1024 x 1024 = 2.21e9 elements each (8.9 GB; a B200 holds arrays far beyond the
    range of 32-bit element offsets): adjoint identity, and the oracle on the LAST slices / rows / views, whose element
    offsets lie above 2^31 (byte offsets above 2^33)."""
    import torch

    n0, n, V = 2112, 1024, 1024
    N, D = (n0, n, n), (n0, n)
    assert n0 * n * n > 2**31 and V * n0 * n > 2**31
    M = _x_mats(N, D, V)
    A = sb.XRayTransform3D(N, M, D)
    assert A.plan_info()["path_name"] == "3d_sep"
    g = torch.Generator(device=cuda_device).manual_seed(7)
    x = torch.randn(N, device=cuda_device, generator=g)
    y = torch.randn(A.output_shape, device=cuda_device, generator=g)
    Ax = A(x)
    ATy = A.adj(y)
    ns, ref = _dots(torch, Ax, y, x, ATy)
    assert ns < TOL and ref < 1e-4, (ns, ref)
    # forward: complete views (the first and the last ones) of the first and the last 4-slice slab
    vs = [0, 300, V - 2, V - 1]
    for z0 in (0, n0 - 4):
        z1 = z0 + 4
        want_f = C.project_3d(x[z0:z1].cpu().numpy(), A.matrices[vs], D, slice_offset=z0, fused=True)[:, z0:z1]
        got_f = Ax[vs][:, z0:z1].cpu().numpy()
        assert max(O.rel_l2(got_f[i], want_f[i]) for i in range(len(vs))) <= TOL, z0
    del Ax, x
    # back projection: voxels of the last 64 slices (and a few of the first ones) against the oracle
    rng = np.random.default_rng(2)
    hi = np.stack([rng.integers(n0 - 64, n0, 1024), rng.integers(0, n, 1024), rng.integers(0, n, 1024)], 1)
    lo = np.stack([rng.integers(0, 8, 64), rng.integers(0, n, 64), rng.integers(0, n, 64)], 1)
    pts = np.concatenate([hi, lo, [[n0 - 1, n - 1, n - 1]]]).astype(np.int32)
    want = C.back_project_3d_points(y.cpu().numpy(), A.matrices, pts)
    got = ATy[pts[:, 0], pts[:, 1], pts[:, 2]].cpu().numpy()
    assert O.rel_l2(got, want) <= TOL


def test_more_than_2_to_the_31_voxels_general_matrices(cuda_device):
    """The brick kernels (tilted geometry: general 2 x 4 matrices) on a 1312^3 volume (2.26e9 voxels, 9.0 GB): adjoint
    identity, and the back projection against the oracle at voxels whose element offsets lie above 2^31."""
    import torch

    n, V = 1312, 6
    N, D = (n, n, n), (1856, 1856)
    assert n**3 > 2**31
    ang = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, np.deg2rad(74.0))], 1)
    A = sb.XRayTransform3D(N, sb.matrices_from_euler_angles(N, D, "XY", ang), D)
    assert A.plan_info()["path_name"] == "3d_general"
    g = torch.Generator(device=cuda_device).manual_seed(8)
    x = torch.randn(N, device=cuda_device, generator=g)
    y = torch.randn(A.output_shape, device=cuda_device, generator=g)
    Ax, ATy = A(x), A.adj(y)
    ns, ref = _dots(torch, Ax, y, x, ATy)
    assert ns < TOL and ref < 1e-4, (ns, ref)
    rng = np.random.default_rng(3)
    hi = np.stack([rng.integers(n - 32, n, 2048), rng.integers(0, n, 2048), rng.integers(0, n, 2048)], 1)
    lo = np.stack([rng.integers(0, n, 512), rng.integers(0, n, 512), rng.integers(0, n, 512)], 1)
    pts = np.concatenate([hi, lo, [[n - 1, n - 1, n - 1], [0, 0, 0]]]).astype(np.int32)
    want = C.back_project_3d_points(y.cpu().numpy(), A.matrices, pts)
    got = ATy[pts[:, 0], pts[:, 1], pts[:, 2]].cpu().numpy()
    assert O.rel_l2(got, want) <= TOL


def test_more_than_2_to_the_31_pixels_2d(cuda_device):
    """2D image of 47104^2 = 2.22e9 pixels (8.9 GB), two views: adjoint identity, both views against the oracle (forward)
    and the first / last image rows of the oracle's back projection (pixel offsets above 2^31)."""
    import torch

    n = 47104
    assert n * n > 2**31
    angles = np.array([0.3, 1.9])
    A = sb.XRayTransform2D((n, n), angles)
    g = torch.Generator(device=cuda_device).manual_seed(9)
    x = torch.randn((n, n), device=cuda_device, generator=g)
    y = torch.randn(A.output_shape, device=cuda_device, generator=g)
    Ax, ATy = A(x), A.adj(y)
    ns, ref = _dots(torch, Ax, y, x, ATy)
    assert ns < TOL, (ns, ref)
    T = A.view_table
    want = C.project_2d(x.cpu().numpy(), T, A.ny, fused=True)
    assert O.rel_l2(Ax.cpu().numpy(), want) <= TOL
    del x, want
    want_a = C.back_project_2d(y.cpu().numpy(), T, (n, n))
    for sl in (slice(0, 64), slice(n - 64, n)):
        assert O.rel_l2(ATy[sl].cpu().numpy(), want_a[sl]) <= TOL, sl
