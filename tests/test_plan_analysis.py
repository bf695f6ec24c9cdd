"""Host-side plan decisions (CPU only): which kernel family a geometry gets, how its views are split
into launch classes, when the TMA / joint-column / two-bin paths apply.  ``xct*_plan_analyse`` runs the
same code as plan creation without touching a CUDA device."""
import warnings

import numpy as np
import pytest

import scico_b200 as sb
from scico_b200 import _lib

PI = np.pi


def _x(N, D, V, turn=PI, **kw):
    return sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, turn, V, endpoint=False)[:, None], **kw)


def test_headline_geometry_takes_the_walk_joint_tma_path():
    n, V = 1024, 1024
    a = sb.XRayTransform3D((n,) * 3, _x((n,) * 3, (n, n), V), (n, n)).analyse()
    assert a["path_name"] == "3d_sep" and a["fwd_kernel"] == 2 and a["adj_kernel"] == 2
    assert a["fwd_joint"] == 1 and a["adj_tma"] == 1 and a["rows_unit"] == 1 and a["rows_consecutive"] == 1
    assert a["fwd_tile"] == 1  # joint forward on the CTA-shared tile
    assert a["adj_interleaved"] == 1  # adjoint reads the slice-interleaved copy of the sinogram
    n_ = sb.XRayTransform3D((n,) * 3, _x((n,) * 3, (n, n), V), (n, n), _flags=_lib.FLAG_NO_TILE).analyse()
    assert n_["fwd_tile"] == 0 and n_["fwd_joint"] == 1
    assert sum(a["joint_views"]) == V and a["two_bin_views"] == [0, 0, 0, 0] and a["fwd_cold"] == 0
    assert sum(1 for c in a["joint_views"] if c) == 4  # half a turn: four (major axis, signs) classes
    assert 0 < a["adj_jump_views"] < V // 16           # only the views next to an axis
    assert a["updates"] == n ** 3 * V


def test_full_turn_has_eight_classes_and_axis_views_do_not_get_their_own():
    N, D, V = (16, 96, 96), (16, 144), 64
    a = sb.XRayTransform3D(N, _x(N, D, V, turn=2 * PI), D).analyse()
    assert sum(1 for c in a["joint_views"] if c) == 8 and sum(a["joint_views"]) == V
    # angles 0 and pi/2 have a zero minor coefficient: they join a neighbouring class
    b = sb.XRayTransform3D(N, sb.matrices_from_euler_angles(N, D, "X", np.array([[0.0], [0.2], [PI / 2], [1.2]])), D).analyse()
    assert sum(b["joint_views"]) == 4 and sum(1 for c in b["joint_views"] if c) == 2


def test_slab_geometry_keeps_consecutive_rows():
    N, D, V = (64, 40, 48), (64, 64), 9
    M = _x(N, D, V)
    a = sb.XRayTransform3D((20,) + N[1:], M, (20, D[1]), slice_offset=30, det_row_offset=30, det_rows_total=64).analyse()
    assert a["rows_unit"] == 1 and a["rows_consecutive"] == 1 and a["adj_tma"] == 1 and a["fwd_joint"] == 1


@pytest.mark.parametrize("case,expect", [
    ("tilt", dict(path_name="3d_general", fwd_kernel=3, adj_kernel=3, fwd_joint=0, adj_tma=1)),  # brick kernels
    ("odd_columns", dict(path_name="3d_sep", fwd_kernel=2, adj_kernel=1, fwd_joint=0, adj_tma=0)),
    ("wide_voxels", dict(path_name="3d_sep", fwd_cold=1, fwd_joint=0)),
    ("row_mixing", dict(path_name="3d_sep", rows_unit=0, fwd_kernel=2, fwd_joint=1, adj_kernel=1, adj_tma=0)),
])
def test_geometries_outside_the_envelope_fall_back(case, expect):
    N, D, V = (12, 40, 44), (12, 96), 16
    if case == "tilt":
        M = sb.matrices_from_euler_angles(N, D, "XY", np.stack([np.linspace(0, PI, V, endpoint=False), np.full(V, 0.5)], 1))
    elif case == "odd_columns":
        D = (12, 97)
        M = _x(N, D, V)
    elif case == "wide_voxels":
        M = _x(N, D, V, voxel_spacing=[1.0, 1.4, 1.4])
    else:
        M = _x(N, D, V, det_spacing=[0.75, 1.0])
    a = sb.XRayTransform3D(N, M, D).analyse()
    for k, v in expect.items():
        assert a[k] == v, (k, a[k], v)


def test_flags_select_the_older_kernel_families():
    N, D, V = (16, 40, 48), (16, 64), 9
    M = _x(N, D, V)
    base = sb.XRayTransform3D(N, M, D).analyse()
    assert (base["fwd_kernel"], base["adj_kernel"], base["fwd_joint"], base["adj_tma"]) == (2, 2, 1, 1)
    a = sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_NO_JOINT).analyse()
    assert a["fwd_joint"] == 0 and a["fwd_kernel"] == 2 and sum(a["two_bin_views"]) == V
    assert sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_NO_TMA).analyse()["adj_tma"] == 0
    a = sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_NO_WALK).analyse()
    assert (a["fwd_kernel"], a["adj_kernel"], a["fwd_joint"]) == (1, 1, 0)
    a = sb.XRayTransform3D(N, M, D, _flags=_lib.FLAG_FORCE_GENERAL).analyse()
    assert a["path_name"] == "3d_general" and (a["fwd_kernel"], a["adj_kernel"]) == (0, 0)


def test_2d_plans():
    a = sb.XRayTransform2D((512, 512), np.linspace(0, PI, 360, endpoint=False)).analyse()
    assert a["path_name"] == "2d_plane" and a["fwd_joint"] == 1 and a["fwd_kernel"] == 2 and a["adj_kernel"] == 1
    assert sum(a["joint_views"]) == 360 and a["two_bin_views"] == [0, 0, 0, 0]
    assert sum(1 for c in a["joint_views"] if c) == 4
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # projected pixel wider than a bin
        w = sb.XRayTransform2D((48, 48), np.linspace(0, PI, 8, endpoint=False), dx=1.2).analyse()
    assert w["fwd_joint"] == 0 and w["fwd_kernel"] == 1
    g = sb.XRayTransform2D((48, 48), np.linspace(0, PI, 8, endpoint=False), _flags=_lib.FLAG_FORCE_GENERAL).analyse()
    assert g["path_name"] == "2d_general"


def _tilt(N, D, V, tilt_deg=74.0, **kw):
    ang = np.stack([np.linspace(0, PI, V, endpoint=False), np.full(V, np.deg2rad(tilt_deg))], 1)
    return sb.matrices_from_euler_angles(N, D, "XY", ang, **kw)


def test_general_matrices_take_the_brick_kernels():
    """The tilted geometry of examples/scripts/ct_projector_comparison_3d.py:44-51 (and of bench.py's view-block
    section): brick kernels in both directions, TMA-staged window, every view in a race-free lattice class."""
    n, V = 256, 64
    D = (n + 64, n + 64)
    a = sb.XRayTransform3D((n,) * 3, _tilt((n,) * 3, D, V), D).analyse()
    assert a["path_name"] == "3d_general" and a["fwd_kernel"] == 3 and a["adj_kernel"] == 3 and a["adj_tma"] == 1
    assert sum(a["brick_views"]) == V
    assert sum(a["brick_views"][1::2]) == 0  # no view needs shared-memory atomics
    # an odd detector width cannot be described by a tensor map (row pitch not a multiple of 16 bytes): cp.async staging
    b = sb.XRayTransform3D((40,) * 3, _tilt((40,) * 3, (60, 61), 8), (60, 61)).analyse()
    assert b["adj_kernel"] == 3 and b["adj_tma"] == 0
    # XCT_FLAG_NO_BRICK / XCT_FLAG_FORCE_GENERAL keep the thread-per-voxel family
    for flag in (_lib.FLAG_NO_BRICK, _lib.FLAG_FORCE_GENERAL):
        c = sb.XRayTransform3D((40,) * 3, _tilt((40,) * 3, (60, 64), 8), (60, 64), _flags=flag).analyse()
        assert c["fwd_kernel"] == 0 and c["adj_kernel"] == 0 and sum(c["brick_views"]) == 0


def test_brick_forward_classes_and_fallbacks():
    N, D = (32, 32, 32), (48, 48)
    # fine voxels: lattice lanes 2 voxels apart land less than one bin apart -> atomic classes only
    fine = sb.XRayTransform3D(N, _tilt(N, D, 6, voxel_spacing=[0.45] * 3), D).analyse()
    assert fine["fwd_kernel"] == 3 and sum(fine["brick_views"][0::2]) == 0 and sum(fine["brick_views"]) == 6
    # coarse voxels: the projected brick does not fit the shared-memory window -> thread-per-voxel kernels
    coarse = sb.XRayTransform3D(N, _tilt(N, D, 6, voxel_spacing=[2.6] * 3), D).analyse()
    assert coarse["fwd_kernel"] == 0 and coarse["adj_kernel"] == 0
    # random orientations: more than one depth class
    rng = np.random.default_rng(5)
    M = sb.matrices_from_euler_angles(N, D, "XYZ", rng.uniform(0, 2 * PI, size=(48, 3)))
    r = sb.XRayTransform3D(N, M, D).analyse()
    assert sum(r["brick_views"]) == 48 and sum(1 for k in (0, 2, 4) if r["brick_views"][k] + r["brick_views"][k + 1]) >= 2
    # a separable geometry never reaches the brick analysis
    sep = sb.XRayTransform3D(N, _x(N, D, 8), D).analyse()
    assert sep["path_name"] == "3d_sep" and sum(sep["brick_views"]) == 0
