"""CPU: Oracle-R against the reference's OWN code. oracle/_ref/libref.so is rs_tracker/align/src/align_icp.cpp
and rs_tracker/common/src/point_cloud_utils.cpp compiled UNMODIFIED from /root/reference (oracle/Makefile `ref`)
against stand-in headers for the absent third-party libraries (oracle/shim). This pins the restatement's
control flow and quirks (mu schedule, unweighted centroids, literal reflection patch, pre-update cost,
quaternion round trip, first-point-per-voxel) to the literal source; the third-party arithmetic inside
(SVD, k-d tree, eigen-solver) is stand-in code on both sides."""
import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O
from realsensetracker_b200 import synth

GOLD = np.load(ROOT / "tests" / "golden" / "oracle_r.npz")
GN = np.load(ROOT / "tests" / "golden" / "oracle_n_160x120.npz")

pytestmark = pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref/libref.so not built (needs /root/reference)")


def test_align_icp3d_matches_the_compiled_reference_bit_for_bit():
    src, dst = GOLD["src"], GOLD["dst"]
    for iters, T0 in ((1, None), (7, None), (9, GOLD["T_small"]), (128, None)):
        ok_r, T_r = O.ref_align_icp3d(src, dst, iters, T0=T0)
        ok_o, T_o = O.align_icp3d(src, dst, iters, T0=T0)
        assert ok_r == ok_o
        assert np.array_equal(T_r, T_o), f"{iters} iterations: max diff {np.abs(T_r - T_o).max()}"
    assert O.ref_align_icp3d(src[:2], dst, 3)[0] is False            # align_icp.cpp:77-79


def test_depth_derived_clouds_full_run():
    intr = tuple(GN["intr"])
    s = O.downsample_voxel(O.remove_nans(O.backproject(GN["frames"][1], intr)), 0.05)
    d = O.downsample_voxel(O.remove_nans(O.backproject(GN["frames"][0], intr)), 0.05)
    ok_r, T_r = O.ref_align_icp3d(s, d, 128)
    ok_o, T_o = O.align_icp3d(s, d, 128)
    assert ok_r and ok_o
    dt, dr = synth.pose_error(T_r, T_o)
    assert dt < 1e-6 and dr < 1e-6, (dt, dr)                           # k-d tree tie order is the only freedom


def test_solve_kabsch_matches():
    src, dst = GOLD["src"], GOLD["dst_big"]
    pairs = np.stack([np.arange(len(src)), np.arange(len(src))], 1)
    for w in (None, GOLD["kabsch_w"]):
        ok_r, T_r = O.ref_solve_kabsch(src, dst, pairs, w)
        ok_o, T_o = O.solve_kabsch(src, dst, pairs, w)
        assert ok_r and ok_o and np.array_equal(T_r, T_o)
    assert O.ref_solve_kabsch(src[:2], dst[:2], pairs[:2])[0] is False  # align_icp.cpp:23-25


def test_cloud_utilities_match():
    src = GOLD["src"]
    assert np.array_equal(O.ref_centroid(src), O.centroid(src))
    bad = src.copy(); bad[3, 0] = np.nan; bad[10, 2] = np.inf
    assert np.array_equal(O.ref_remove_nans(bad), O.remove_nans(bad))
    a, b = O.ref_downsample_voxel(src, 0.25), O.downsample_voxel(src, 0.25)
    key = lambda x: x[np.lexsort(x.T[::-1])]
    assert a.shape == b.shape and np.array_equal(key(a), key(b))       # same points; order is the documented deviation
    idx_r, d2_r = O.ref_find_correspondences(GOLD["dst"], src)
    idx_o, d2_o = O.nn(GOLD["dst"], src)
    assert np.array_equal(idx_r, idx_o) and np.array_equal(d2_r, d2_o)


def test_extents_and_orient_normals_match():
    """ComputeExtents (point_cloud_utils.cpp:26-32) and OrientNormals (:205-216) through the reference's own code vs the
    numpy restatements: bit for bit."""
    rng = np.random.default_rng(3)
    for cloud in (GOLD["src"], GOLD["dst"][:1], (rng.normal(size=(5000, 3)) * 3).astype(np.float32)):
        (lo_r, hi_r), (lo_o, hi_o) = O.ref_extents(cloud), O.extents(cloud)
        assert np.array_equal(lo_r, lo_o) and np.array_equal(hi_r, hi_o)
        nrm = rng.normal(size=cloud.shape).astype(np.float32)
        nrm[0] = 0.0
        for vp in ((0.0, 0.0, 0.0), (0.3, -0.2, 1.5)):
            got = O.ref_orient_normals(cloud, vp, nrm)
            assert np.array_equal(got, O.orient_normals(cloud, vp, nrm))
            assert (((cloud.astype(np.float64) - np.asarray(vp)) * got).sum(1) <= 1e-5).all()
    lo, hi = O.extents(np.zeros((0, 3), np.float32))
    assert (lo > 3e38).all() and (hi < -3e38).all()


def test_reference_normals_follow_the_orientation_rule():
    """ComputeNormals + OrientNormals (point_cloud_utils.cpp:176-216): the convention the CUDA normal
    kernel inherits — n . (p - viewpoint) <= 0 — on a plane seen from the origin."""
    rng = np.random.default_rng(0)
    xy = rng.uniform(-1, 1, size=(400, 2))
    n = np.array([0.3, -0.2, -1.0]); n /= np.linalg.norm(n)
    z = (-2.0 - xy @ n[:2]) / n[2]
    pts = np.column_stack([xy, z]).astype(np.float32)
    nr = O.ref_normals(pts, k=16)
    assert np.allclose(np.abs(nr @ n), 1.0, atol=1e-3)
    assert ((nr * pts).sum(1) <= 0).all()


def test_depth_pipeline_through_the_reference_functions():
    """back-project -> RemoveNans -> DownsampleVoxel -> AlignIcp3d through the reference's own functions vs Oracle-R:
    same algorithm, different cloud order after DownsampleVoxel (unordered_map iteration order vs first occurrence),
    so poses agree to rounding, not bits."""
    f, intr = GN["frames"], tuple(GN["intr"])
    ok_r, T_r = O.ref_align_depth_pairs(f[1:3], f[0:2], intr, n_threads=2)
    ok_o, T_o = O.align_depth_pairs(f[1:3], f[0:2], intr, n_threads=2)
    assert ok_r.all() and ok_o.all()
    for i in range(2):
        dt, dr = synth.pose_error(T_r[i], T_o[i])
        assert dt < 1e-4 and dr < 1e-4, (dt, dr)
