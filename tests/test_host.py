"""Host-side logic: geometry tables, operator protocol, error behaviour, C-ABI surface (no GPU)."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

import scico_b200 as sb
from scico_b200 import _lib, geometry
from oracle import xray_np as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- geometry -----------------------------------------------------------------------------
@pytest.mark.parametrize("nx,V,dx", [((12, 13), 10, None), ((64, 64), 90, 0.5), ((300, 200), 50, (0.6, 0.7))])
def test_view_table_matches_oracle_bitwise(nx, V, dx):
    angles = np.linspace(0, np.pi, V, endpoint=False)
    A = sb.XRayTransform2D(nx, angles, dx=dx)
    T = O.view_table_2d(angles, A.x0, A.dx, A.y0)
    np.testing.assert_array_equal(A.view_table, T)
    assert A.view_table.dtype == np.float32 and A.view_table.flags.c_contiguous


def test_euler_matches_scipy_and_oracle():
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(2)
    for seq in ["X", "Y", "Z", "XY", "xy", "XYZ", "zyx"]:
        ang = rng.uniform(-3, 3, (5, len(seq)))
        np.testing.assert_allclose(geometry.euler_matrices(seq, ang), Rotation.from_euler(seq, ang).as_matrix(), atol=1e-14)
        np.testing.assert_allclose(
            sb.matrices_from_euler_angles((5, 6, 7), (8, 9), seq, ang, voxel_spacing=[1, 2, 3], det_spacing=[0.5, 2]),
            O.matrices_from_euler_angles((5, 6, 7), (8, 9), seq, ang, voxel_spacing=[1, 2, 3], det_spacing=[0.5, 2]),
            atol=1e-13,
        )
    with pytest.raises(ValueError):
        geometry.euler_matrices("Xy", np.zeros((2, 2)))


def test_x_rotation_is_separable_with_exact_zeros():
    M = sb.matrices_from_euler_angles((16,) * 3, (16, 16), "X", np.linspace(0, np.pi, 7, endpoint=False)[:, None])
    assert geometry.is_axis0_separable(M)
    M2 = sb.matrices_from_euler_angles((16,) * 3, (16, 16), "XY", np.array([[0.3, 1.2]]))
    assert not geometry.is_axis0_separable(M2)


def test_slab_row_range():
    M = sb.matrices_from_euler_angles((64, 8, 8), (64, 12), "X", np.linspace(0, np.pi, 5, endpoint=False)[:, None])
    assert geometry.slab_row_range(M, 16, 32, 64) == (16, 32)
    M = sb.matrices_from_euler_angles((64, 8, 8), (80, 12), "X", np.zeros((1, 1)))  # t0 = 8
    assert geometry.slab_row_range(M, 0, 16, 80) == (8, 24)


# ---- operator protocol ----------------------------------------------------------------------
@pytest.mark.filterwarnings("error")
def test_init_warnings():
    """scico/test/linop/xray/test_xray_2d.py:14-30."""
    sb.XRayTransform2D((3, 3), np.array([np.pi / 4]))
    sb.XRayTransform2D((3, 3), np.array([0]), dx=np.array([1, 1]))
    with pytest.warns(UserWarning):
        sb.XRayTransform2D((3, 3), np.array([0.1]), dx=np.array([1.1, 1.1]))
    with pytest.warns(UserWarning):
        sb.XRayTransform2D((3, 3), np.array([0]), dx=np.array([1.1, 1.1]))


def test_2d_attributes_and_defaults():
    angles = np.linspace(0, np.pi, 10, endpoint=False)
    A = sb.XRayTransform2D((12, 13), angles)
    assert A.input_shape == (12, 13) and A.output_shape == (10, 18)
    assert A.det_count == A.ny == 18 and A.y0 == -9.0 and A.dy == 1.0
    assert A.input_dtype == np.float32 and A.output_dtype == np.float32
    assert A.input_size == 156 and A.output_size == 180 and A.matrix_shape == (180, 156)
    assert A.shape == ((10, 18), (12, 13))
    np.testing.assert_allclose(A.x0, -(np.array([12, 13]) * np.sqrt(2) / 2) / 2)
    assert sb.XRayTransform2D((12, 13), angles, det_count=14).output_shape == (10, 14)
    assert sb.XRayTransform2D((12, 13), angles, dx=0.5).dx == (0.5, 0.5)


def test_3d_attributes():
    M = sb.matrices_from_euler_angles((4, 5, 6), (7, 8), "X", np.zeros((3, 1)))
    A = sb.XRayTransform3D((4, 5, 6), M, [7, 8])
    assert A.output_shape == (3, 7, 8) and A.det_shape == (7, 8) and A.batch_size == 8
    assert A.matrices.dtype == np.float32 and A.matrices.shape == (3, 2, 4)
    assert hasattr(sb.XRayTransform3D, "matrices_from_euler_angles")
    with pytest.raises(ValueError):
        sb.XRayTransform3D((4, 5, 6), np.zeros((3, 2, 3)), (7, 8))


def test_shape_and_dtype_errors():
    """Operator.__call__: ValueError on shape mismatch, no dtype check (_operator.py:219-223);
    LinearOperator.adj: ValueError on dtype or shape mismatch (_linop.py:319-325)."""
    A = sb.XRayTransform2D((12, 13), np.linspace(0, np.pi, 10, endpoint=False))
    with pytest.raises(ValueError):
        A(np.zeros((13, 12), np.float32))
    with pytest.raises(ValueError):
        A.adj(np.zeros((10, 18), np.float64))
    with pytest.raises(ValueError):
        A.adj(np.zeros((10, 17), np.float32))
    with pytest.raises(ValueError):
        A.T(np.zeros((10, 17), np.float32))


def test_transpose_wiring_with_fake_kernels():
    """.T / .H / .adj / gram_op / @ / scaling resolve to the two kernels (_linop.py:296-440)."""
    Mx = np.arange(12, dtype=np.float32).reshape(3, 4)
    calls = []
    A = sb.LinearOperator((4,), (3,), eval_fn=lambda x: (calls.append("f"), Mx @ x)[1],
                          adj_fn=lambda y: (calls.append("a"), Mx.T @ y)[1])
    x, y = np.ones(4, np.float32), np.ones(3, np.float32)
    np.testing.assert_array_equal(A @ x, Mx @ x)
    np.testing.assert_array_equal(A.T @ y, Mx.T @ y)
    np.testing.assert_array_equal(A.H(y), Mx.T @ y)
    np.testing.assert_array_equal(A.conj().T(y), Mx.T @ y)
    np.testing.assert_array_equal(A.T.T(x), Mx @ x)
    np.testing.assert_array_equal(A.gram_op(x), Mx.T @ (Mx @ x))
    np.testing.assert_array_equal((2.0 * A)(x), 2 * (Mx @ x))
    np.testing.assert_array_equal((2.0 * A).adj(y), 2 * (Mx.T @ y))
    np.testing.assert_array_equal((A + A)(x), 2 * (Mx @ x))
    np.testing.assert_array_equal((A.T @ A)(x), Mx.T @ (Mx @ x))
    assert A.T.input_shape == (3,) and A.T.output_shape == (4,)
    assert sb.valid_adjoint(A, A.T, eps=1e-6)
    assert abs(sb.operator_norm(A, maxiter=200) - np.linalg.norm(Mx, 2)) < 1e-3
    with pytest.raises(TypeError):
        sb.LinearOperator((4,), (3,), eval_fn=lambda x: x, adj_fn=3)


def test_autograd_resolves_to_the_adjoint_kernel(monkeypatch):
    """torch.autograd through the pair (the reference's custom_vjp wiring for external projectors,
    _astra_3d.py:498-502; checks of test/linop/xray/astra/test_astra_2d.py:145-187): with the two kernels
    replaced by a dense matrix, grad ||A x||^2 = 2 A^T A x, the gradient through A.T, a second derivative,
    and no tape (and no detour) for inputs that do not require grad."""
    import torch

    from scico_b200 import xray

    nx, V = (5, 4), 3
    A = sb.XRayTransform2D(nx, np.linspace(0, np.pi, V, endpoint=False))
    rng = np.random.default_rng(0)
    M = torch.from_numpy(rng.standard_normal((V * A.ny, nx[0] * nx[1])).astype(np.float32))
    calls = []

    def fake_apply(plans, x, out_shape, forward, batch, default_device, out=None, wait=True):
        assert plans is A._plans
        assert out is not None or not (torch.is_grad_enabled() and x.requires_grad)  # kernels run off the tape
        x = x.detach()
        calls.append("f" if forward else "a")
        return ((M if forward else M.T) @ x.reshape(-1)).reshape(tuple(out_shape))

    monkeypatch.setattr(xray, "_apply", fake_apply)
    x = torch.from_numpy(rng.standard_normal(nx).astype(np.float32)).requires_grad_()
    y = torch.from_numpy(rng.standard_normal(A.output_shape).astype(np.float32)).requires_grad_()

    (g,) = torch.autograd.grad((A(x) ** 2).sum(), x)
    want = 2 * (M.T @ (M @ x.detach().reshape(-1))).reshape(nx)
    torch.testing.assert_close(g, want, rtol=1e-5, atol=1e-5)
    assert calls == ["f", "a"]  # the backward pass IS the back projection kernel

    calls.clear()
    (gy,) = torch.autograd.grad((A.T(y) ** 2).sum(), y)  # gradient through the transpose
    torch.testing.assert_close(gy, 2 * (M @ (M.T @ y.detach().reshape(-1))).reshape(A.output_shape))
    assert calls == ["a", "f"]

    # a second derivative: d/dx <grad_x 1/2||Ax||^2, c> = A^T A c
    c = torch.from_numpy(rng.standard_normal(nx).astype(np.float32))
    (g1,) = torch.autograd.grad(0.5 * (A(x) ** 2).sum(), x, create_graph=True)
    (g2,) = torch.autograd.grad((g1 * c).sum(), x)
    torch.testing.assert_close(g2, (M.T @ (M @ c.reshape(-1))).reshape(nx))

    # composite operators built by the LinearOperator algebra stay differentiable
    (g3,) = torch.autograd.grad(((2.0 * A).gram_op(x) * c).sum(), x)
    torch.testing.assert_close(g3, 4 * (M.T @ (M @ c.reshape(-1))).reshape(nx))

    calls.clear()
    with torch.no_grad():
        assert not A(x).requires_grad
    assert not A(x.detach()).requires_grad and calls == ["f", "f"]
    out = torch.empty(A.output_shape, dtype=torch.float32)
    with pytest.raises(ValueError):  # an explicit result buffer cannot be recorded on the tape: refused, not detached
        A.project(x, out=out)
    with torch.no_grad():
        assert not A.project(x, out=out).requires_grad


# ---- C ABI ----------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "scico_b200_xray.h")).read()
    declared = set(re.findall(r"XCT_API\s+[\w\s\*]+?\b(xct\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.xct_version() == 100


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.Geom2D) == 32 and ctypes.sizeof(_lib.Geom3D) == 56
    assert ctypes.sizeof(_lib.PlanInfo) == 64


def test_route_structs_match_the_header_as_gcc_lays_them_out(tmp_path):
    """xct_out_route / xct_ipc_handle: ctypes mirror against sizeof / offsetof of the C header."""
    import subprocess

    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "scico_b200_xray.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %d\\n", sizeof(xct_out_route), offsetof(xct_out_route, nparts),'
        ' offsetof(xct_out_route, row_begin), offsetof(xct_out_route, ptr), offsetof(xct_out_route, store),'
        ' sizeof(xct_ipc_handle), XCT_MAX_ROUTE_PARTS); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(t) for t in subprocess.check_output([str(exe)]).split()]
    R = _lib.OutRoute
    assert got == [ctypes.sizeof(R), R.nparts.offset, R.row_begin.offset, R.ptr.offset, R.store.offset,
                   ctypes.sizeof(_lib.IpcHandle), _lib.MAX_ROUTE_PARTS]


def test_scatter_and_peer_entry_points_report_errors_without_a_device():
    import torch

    L = _lib.lib()
    assert L.xct_adjoint_scatter(None, None, None, None) == _lib.XCT_ERR_INVALID
    assert L.xct_sum_slots(0, None, None, 1, 4, 4, None) == _lib.XCT_ERR_INVALID
    assert L.xct_peer_zero(0, None, 16, None) == _lib.XCT_ERR_INVALID
    assert L.xct_peer_copy_out(0, None, None, 16, None) == _lib.XCT_ERR_INVALID
    assert L.xct_peer_close(0, None) == 0 and L.xct_peer_free(0, None) == 0  # no-ops
    ptr, h = ctypes.c_void_p(), _lib.IpcHandle()
    assert L.xct_peer_alloc(0, 0, ctypes.byref(ptr), ctypes.byref(h)) == _lib.XCT_ERR_INVALID
    A = sb.XRayTransform2D((12, 13), np.linspace(0, np.pi, 10, endpoint=False))
    with pytest.raises(ValueError):  # no routed path for host arrays
        A.back_project_scatter(np.zeros(A.output_shape, np.float32), [1], [0, 12])
    if not torch.cuda.is_available():
        assert L.xct_peer_alloc(0, 1024, ctypes.byref(ptr), ctypes.byref(h)) == _lib.XCT_ERR_NO_DEVICE
        assert L.xct_peer_open(0, ctypes.byref(h), ctypes.byref(ptr)) == _lib.XCT_ERR_NO_DEVICE
        from scico_b200 import sharded

        with pytest.raises(RuntimeError):
            sharded.PeerBlocks([(0, 12)], (13,), rank=0, world_size=1)
    with pytest.raises(ValueError):
        from scico_b200 import sharded

        sharded.ViewShardedXRayTransform2D((12, 13), np.linspace(0, np.pi, 10, endpoint=False), rank=0, world_size=1,
                                           exchange="carrier pigeon")


def test_native_peer_memory_calls_the_c_abi_with_well_typed_arguments(monkeypatch):
    """sharded._NativePeerMemory (the product's backend of PeerBlocks) against a recording double of the
    library: every call must convert under the real ctypes argtypes, and PeerBlocks must drive it in the
    documented order (alloc + zero per copy, sum_slots / copy_out + zero per exchange, free on close)."""
    import torch

    from scico_b200 import sharded

    real = _lib.lib()
    calls = []

    class Fake:
        def __getattr__(self, name):
            fn = getattr(real, name)

            def call(*args):
                assert len(args) == len(fn.argtypes), (name, args)
                for a, t in zip(args, fn.argtypes):
                    t.from_param(a)  # raises ctypes.ArgumentError on a badly typed argument
                calls.append(name)
                if name == "xct_peer_alloc":
                    args[2]._obj.value = 0x7000_0000_0000 + 0x1000_0000 * calls.count("xct_peer_alloc")
                return 0

            return call

    class Stream:
        cuda_stream = 0

    monkeypatch.setattr(_lib, "lib", lambda: Fake())
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: Stream())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(sharded._NativePeerMemory, "token", lambda self: torch.zeros(1))
    for mode, per_call in (("store", ["xct_sum_slots"]), ("add", ["xct_peer_copy_out", "xct_peer_zero"])):
        calls.clear()
        pb = sharded.PeerBlocks([(0, 6)], (5,), rank=0, world_size=1, mode=mode)
        assert calls == ["xct_peer_alloc", "xct_peer_zero"] * 3  # two block copies + the rendezvous flag words
        assert pb.ptrs[0][0] != pb.ptrs[1][0] and pb.local_shape == (6, 5) and pb.row_begin == [0, 6]
        seen = []
        out = torch.empty((6, 5))
        for _ in range(3):
            pb.exchange(lambda ptrs, rb, store: seen.append((ptrs[0], store)), out)
        assert [p for p, _ in seen] == [pb.ptrs[0][0], pb.ptrs[1][0], pb.ptrs[0][0]]  # the copies alternate
        assert all(st == (mode == "store") for _, st in seen)
        assert calls[6:] == per_call * 3  # world size 1: no rendezvous
        pb.close()
        assert calls[-3:] == ["xct_peer_free"] * 3
        pb.close()
    # the flag rendezvous converts under the real argtypes too
    mem = sharded._NativePeerMemory(0)
    monkeypatch.setattr(torch, "zeros", lambda *a, **k: type("T", (), {"data_ptr": lambda self: 0x7100_0000_0000})())
    calls.clear()
    mem.signal([0x7000_0000_0000, 0x7000_0000_0040], 3)
    mem.wait_flags(0x7000_0000_0100, 2, 3, 1.5)
    assert calls == ["xct_peer_signal", "xct_peer_wait"]
    monkeypatch.undo()
    monkeypatch.setattr(_lib, "lib", lambda: Fake())
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    # opening a peer's handle: 64 bytes in, device address out
    mem = sharded._NativePeerMemory(0)
    assert isinstance(mem.open(bytes(range(64))), (int, type(None)))
    assert mem.offset(4096, 128) == 4224


def test_invalid_arguments_are_reported_not_fatal():
    L = _lib.lib()
    pl = ctypes.c_void_p()
    g = _lib.Geom2D()
    assert L.xct2d_plan_create(ctypes.byref(pl), ctypes.byref(g)) == _lib.XCT_ERR_INVALID
    assert b"xct2d_plan_create" in L.xct_last_error()
    assert L.xct_forward(None, None, None, 1, None) == _lib.XCT_ERR_INVALID
    assert L.xct_plan_get_info(None, None) == _lib.XCT_ERR_INVALID
    L.xct_plan_destroy(None)  # no-op


def test_no_cpu_fallback():
    """Without a CUDA device every compute call must fail loudly."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    A = sb.XRayTransform2D((12, 13), np.linspace(0, np.pi, 10, endpoint=False))
    with pytest.raises(_lib.XctError) as ei:
        A(np.ones((12, 13), np.float32))
    assert ei.value.code == _lib.XCT_ERR_NO_DEVICE
    M = sb.matrices_from_euler_angles((4, 4, 4), (4, 4), "X", np.zeros((1, 1)))
    with pytest.raises(_lib.XctError):
        sb.XRayTransform3D((4, 4, 4), M, (4, 4)).adj(np.ones((1, 4, 4), np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "scico_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle's", "").replace("as the oracle", "") or f.endswith((".cuh", ".cu")), f
                if f.endswith(".py"):
                    assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
    # tools/ runs the product alone; bench.py touches the oracle only inside its CPU arm (time_cpu_port)
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            assert not re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(ROOT, "tools", f)).read(), re.M), f
    bench = open(os.path.join(ROOT, "bench.py")).read()
    assert not re.search(r"^(from|import)\s+oracle", bench, re.M)  # no module-level import
    uses = [m.start() for m in re.finditer(r"from oracle import", bench)]
    lo, hi = bench.index("def time_cpu_port"), bench.index("class ClockSampler")
    assert uses and all(lo < u < hi for u in uses)


# ---- ASTRA-free geometry conversion (scico/linop/xray/astra/_astra_3d.py:185-307,595-631) ------------
def _fixture_vectors():
    """scico/test/linop/xray/astra/test_astra_3d.py:123-155: project along z, rows with y, columns
    with x, detector centre shifted by (1, 2, 3)."""
    return np.array([[0, 0, 1, 1, 2, 3, 1, 0, 0, 0, 1, 0]], dtype=np.float64)


def test_convert_to_scico_geometry_known_answer():
    """test_astra_3d.py:200-207: the fixture converts to [[[0,1,0,-2],[0,0,1,-1]]]."""
    from scico_b200 import geometry as G

    M = G.convert_to_scico_geometry((30, 31, 32), (31, 32), vectors=_fixture_vectors())
    np.testing.assert_allclose(M, np.array([[[0.0, 1.0, 0.0, -2.0], [0.0, 0.0, 1.0, -1.0]]]), atol=1e-12)
    # test_astra_3d.py:185-197: voxel (i, j, k) lands on detector index (j - 2, k - 1)
    rng = np.random.default_rng(0)
    ijk = np.array([rng.integers(0, s) for s in (30, 31, 32)], dtype=np.float64)
    v = _fixture_vectors()[0]
    got = G.project_world_coordinates(G.volume_coords_to_world_coords(ijk, (30, 31, 32)), v[0:3], v[3:6], v[6:9], v[9:12], (31, 32))
    np.testing.assert_allclose(got, [ijk[1] - 2, ijk[2] - 1], atol=1e-12)
    with pytest.raises(ValueError):
        G.convert_to_scico_geometry((4, 4, 4), (4, 4), angles=np.zeros(2), vectors=_fixture_vectors())
    with pytest.raises(ValueError):
        G.convert_to_scico_geometry((4, 4, 4), (4, 4))


def test_convert_from_scico_geometry_known_answer_and_round_trip():
    """test_astra_3d.py:210-222 (element 5, the detector centre along the ray, is free)."""
    from scico_b200 import geometry as G

    M = np.array([[[0.0, 1.0, 0.0, -2.0], [0.0, 0.0, 1.0, -1.0]]])
    vec = G.convert_from_scico_geometry((30, 31, 32), M, (31, 32))
    truth = _fixture_vectors()
    np.testing.assert_allclose(vec[0, :5], truth[0, :5], atol=1e-12)
    np.testing.assert_allclose(vec[0, 6:], truth[0, 6:], atol=1e-12)
    # round trip through vectors for a tilted geometry (unit detector spacing: the reference's formula
    # takes the matrix rows themselves as u and v, which inverts the conversion only for unit steps)
    from scipy.spatial.transform import Rotation

    ang = np.linspace(0, np.pi, 7, endpoint=False)
    v0 = G.rotate_vectors(G.angle_to_vector((1.0, 1.0), ang), Rotation.from_euler("x", 0.3))
    M0 = G.convert_to_scico_geometry((12, 14, 16), (20, 24), vectors=v0)
    M1 = G.convert_to_scico_geometry((12, 14, 16), (20, 24), vectors=G.convert_from_scico_geometry((12, 14, 16), M0, (20, 24)))
    np.testing.assert_allclose(M1, M0, atol=1e-10)


def test_angle_to_vector_and_rotate_vectors():
    """test_astra_3d.py:104-118."""
    from scipy.spatial.transform import Rotation

    from scico_b200 import geometry as G

    assert G.angle_to_vector([0.9, 1.5], np.linspace(0, np.pi, 5)).shape == (5, 12)
    v0 = G.angle_to_vector([1.0, 1.0], np.linspace(0, np.pi / 2, 4, endpoint=False))
    v1 = G.angle_to_vector([1.0, 1.0], np.linspace(np.pi / 2, np.pi, 4, endpoint=False))
    r = Rotation.from_euler("z", np.pi / 2)
    np.testing.assert_allclose(G.rotate_vectors(v0, r), v1, atol=1e-7)
    np.testing.assert_allclose(G.rotate_vectors(v0, r.as_matrix()), v1, atol=1e-7)


def test_parallel3d_geometry_is_axis0_separable():
    """ct_3d_tv_padmm.py:49-58: angle_to_vector + convert_to_scico_geometry gives rows that depend on
    volume axis 0 only -- the geometry the z-slab sharding and the walk kernels are built for."""
    from scico_b200 import geometry as G

    shape, det = (64, 128, 256), (64, 256)
    angles = np.linspace(0, np.pi, 10, endpoint=False)
    M = G.convert_to_scico_geometry(shape, det, vectors=G.angle_to_vector([1.0, 1.0], angles))
    M2 = G.convert_to_scico_geometry(shape, det, det_spacing=[1.0, 1.0], angles=angles)
    np.testing.assert_allclose(M, M2, atol=1e-12)
    assert G.is_axis0_separable(M.astype(np.float32))  # exact zeros: no pseudo-inverse noise left
    np.testing.assert_allclose(M[:, 0], np.tile([1.0, 0.0, 0.0, 0.0], (10, 1)), atol=1e-12)


def test_operator_registry_ids_lifetime_and_errors_without_a_device():
    """xct_op_* (what the XLA FFI handlers call): registration validates and COPIES the geometry, ids are never
    reused, retain / release count references, a released id is an error.  Plans need a device: without one
    xct_op_plan reports XCT_ERR_NO_DEVICE (no CPU fallback), and applying without a plan is refused."""
    from scico_b200.jax_ffi import RegisteredOperator

    L = _lib.lib()
    A = sb.XRayTransform2D((12, 10), np.linspace(0, np.pi, 5, endpoint=False))
    r1, r2 = RegisteredOperator(A), RegisteredOperator(A)
    assert r2.id == r1.id + 1 and r1.input_shape == (12, 10) and r1.output_shape == A.output_shape
    assert L.xct_op_retain(r1.id) == 0
    r1.release()                                   # one reference left (the retain above)
    assert L.xct_op_retain(r1.id) == 0 and L.xct_op_release(r1.id) == 0
    assert L.xct_op_release(r1.id) == 0            # last reference
    assert L.xct_op_release(r1.id) == _lib.XCT_ERR_INVALID and L.xct_op_retain(r1.id) == _lib.XCT_ERR_INVALID
    pl = ctypes.c_void_p()
    if L.xct_device_count() == 0:
        assert L.xct_op_plan(r2.id, 0, ctypes.byref(pl)) == _lib.XCT_ERR_NO_DEVICE
    assert L.xct_op_apply(r2.id, 7, 1, None, None, 120, None) == _lib.XCT_ERR_INVALID  # no plan on device 7
    assert b"no plan on this device" in L.xct_last_error()
    r3 = RegisteredOperator(sb.XRayTransform3D((4, 5, 6), sb.matrices_from_euler_angles((4, 5, 6), (7, 8), "X", np.zeros((2, 1))), (7, 8)))
    assert r3.id == r2.id + 1
    g = _lib.Geom3D()                               # invalid geometry: refused at registration
    oid = ctypes.c_int64()
    assert L.xct_op_register_3d(ctypes.byref(g), ctypes.byref(oid)) == _lib.XCT_ERR_INVALID


def test_device_arguments_accept_a_partition_like_the_reference_accepts_a_sharding():
    """input_device / output_device (scico/linop/xray/_xray3d.py:61-62): a device ordinal places the arrays, a
    scico_b200.sharded.Partition makes the constructor return this rank's partitioned operator."""
    from scico_b200 import sharded

    N, D, V = (16, 12, 10), (16, 14), 6
    Mx = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
    Mt = sb.matrices_from_euler_angles(N, D, "XY", np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, 0.4)], 1))
    P = lambda kind, r: sharded.Partition(kind, rank=r, world_size=4)  # noqa: E731
    A = sb.XRayTransform3D(N, Mx, D, input_device=P("auto", 1))
    assert isinstance(A, sharded.SlabShardedXRayTransform3D) and A.slab == (4, 8) and A.rows == (4, 8)
    assert isinstance(A.local, sb.XRayTransform3D) and A.local.slice_offset == 4 and A.local.det_row_offset == 4
    B = sb.XRayTransform3D(N, Mt, D, output_device=P("auto", 3))
    assert isinstance(B, sharded.ViewShardedXRayTransform3D) and B.views == (4, 6) and B.slab == (12, 16)
    C_ = sb.XRayTransform3D(N, Mx, D, input_device=P("views", 0), output_device=P("views", 0))
    assert isinstance(C_, sharded.ViewShardedXRayTransform3D) and C_.local_output_shape == (1,) + D
    with pytest.raises(ValueError):
        sb.XRayTransform3D(N, Mt, D, input_device=P("slabs", 0))       # not separable
    with pytest.raises(ValueError):
        sb.XRayTransform3D(N, Mx, D, input_device=P("slabs", 0), output_device=P("views", 0))
    with pytest.raises(ValueError):
        sb.XRayTransform3D(N, Mx, D, input_device=P("slabs", 0), slice_offset=2)
    E = sb.XRayTransform2D((32, 24), np.linspace(0, np.pi, 8, endpoint=False), det_count=50, output_device=P("views", 2))
    assert isinstance(E, sharded.ViewShardedXRayTransform2D) and E.views == (4, 6) and E.ny == 50
    with pytest.raises(ValueError):
        sb.XRayTransform2D((32, 24), np.linspace(0, np.pi, 8, endpoint=False), input_device=P("slabs", 0))
    plain = sb.XRayTransform3D(N, Mx, D, input_device=0)
    assert type(plain) is sb.XRayTransform3D and plain.input_device == 0
