"""CPU: Oracle-R (restatement of the reference's AlignIcp3d / SolveKabsch / cloud utilities) against
scipy / numpy and the committed golden vectors. Anchors: align_icp.cpp:18-167, kdtree.hpp:51-57,
point_cloud_utils.cpp:34-98,163-174, and the disabled self-test rs_align_app.cpp:257-263."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from conftest import ROOT
from oracle import oracle as O
from realsensetracker_b200 import synth

GOLD = np.load(ROOT / "tests" / "golden" / "oracle_r.npz")


def test_exact_nn_matches_ckdtree():
    rng = np.random.default_rng(0)
    for m, n in ((1, 5), (17, 100), (5000, 3000)):
        dst = rng.uniform(-1, 1, size=(m, 3)).astype(np.float32)
        q = rng.uniform(-1.2, 1.2, size=(n, 3)).astype(np.float32)
        idx, d2 = O.nn(dst, q, leaf=16)
        dd, ii = cKDTree(dst.astype(np.float64)).query(q.astype(np.float64))
        same = idx == ii
        assert same.mean() > 0.999
        assert np.allclose(d2[same], dd[same] ** 2, rtol=1e-5, atol=1e-10)
        # where they differ it is an fp32 near-tie: the chosen neighbour is as close as scipy's
        d_alt = ((dst[idx[~same]].astype(np.float64) - q[~same]) ** 2).sum(1)
        assert np.allclose(d_alt, dd[~same] ** 2, rtol=1e-5)


def test_centroid_is_sequential_fp32():
    pts = GOLD["src"]
    s = np.zeros(3, dtype=np.float32)
    for p in pts:
        s += p
    assert np.array_equal(O.centroid(pts), s * np.float32(1.0 / len(pts)))   # point_cloud_utils.cpp:92-98


def test_remove_nans_and_voxel_downsample():
    pts = GOLD["src"].copy()
    pts[5, 1] = np.nan; pts[9, 0] = np.inf
    out = O.remove_nans(pts)
    assert len(out) == len(pts) - 2 and np.isfinite(out).all()
    assert np.array_equal(out, pts[np.isfinite(pts).all(1)])
    src = GOLD["src"]
    vox = O.downsample_voxel(src, 0.25)
    keys = np.floor(src / np.float32(0.25)).astype(np.int64)
    _, first = np.unique(keys, axis=0, return_index=True)
    assert np.array_equal(vox, src[np.sort(first)])                          # first point per voxel, first-seen order
    assert np.array_equal(vox, GOLD["voxel_025"])


def test_kabsch_closed_form_and_golden():
    src, dst = GOLD["src"], GOLD["dst_big"]
    pairs = np.stack([np.arange(len(src)), np.arange(len(src))], 1)
    ok, T = O.solve_kabsch(src, dst, pairs)
    assert ok and np.array_equal(T, GOLD["kabsch_T"])
    assert synth.pose_error(T, GOLD["T_big"]) < (1e-5, 1e-5)                 # the reference's known rotation
    # independent numpy SVD
    sm, dm = src.mean(0), dst.mean(0)
    U, _, Vt = np.linalg.svd((dst - dm).T.astype(np.float64) @ (src - sm).astype(np.float64))
    assert np.allclose(T[:3, :3], U @ Vt, atol=1e-5)
    okw, Tw = O.solve_kabsch(src, dst, pairs, GOLD["kabsch_w"])
    assert okw and np.array_equal(Tw, GOLD["kabsch_T_weighted"])
    assert synth.pose_error(Tw, GOLD["T_big"]) < (1e-5, 1e-5)
    assert O.solve_kabsch(src[:2], dst[:2], pairs[:2])[0] is False           # < 3 points, align_icp.cpp:23-25


def test_align_icp3d_known_motion_and_golden():
    src, dst = GOLD["src"], GOLD["dst"]
    ok, T, ex = O.align_icp3d(src, dst, 128, details=True)
    assert ok
    assert np.array_equal(T, GOLD["icp_T"])
    assert ex["mean_cost"] == np.float32(GOLD["icp_mean_cost"])
    assert np.array_equal(ex["nbrs"], GOLD["icp_nbrs"]) and np.array_equal(ex["weights"], GOLD["icp_weights"])
    assert np.allclose(ex["cov"], GOLD["icp_cov"], rtol=1e-12)
    assert synth.pose_error(T, GOLD["T_small"]) < (1e-4, 1e-4)
    assert (ex["nbrs"] == np.arange(len(src))).mean() > 0.99                 # converged: every point found itself
    # weights are the Geman-McClure form with the annealed mu (align_icp.cpp:96-98,116-118)
    mu = np.float32(1.0)
    for it in range(128):
        if it > 0 and it % 8 == 0:
            mu = mu / np.float32(1.4)
    assert np.all(ex["weights"] <= 1.0) and ex["weights"].min() > 0.0
    assert O.align_icp3d(src[:2], dst, 5)[0] is False                        # align_icp.cpp:77-79


def test_initial_pose_is_read_and_result_overwrites_it():
    src, dst = GOLD["src"], GOLD["dst"]
    ok0, T0 = O.align_icp3d(src, dst, 0, T0=GOLD["T_small"])
    assert ok0 and np.allclose(T0, GOLD["T_small"], atol=1e-6)               # 0 iterations: pose passes through
    ok1, T1 = O.align_icp3d(src, dst, 1, T0=GOLD["T_small"])
    assert synth.pose_error(T1, GOLD["T_small"]) < (1e-5, 1e-5)


def test_depth_pair_pipeline_matches_golden_and_ground_truth():
    g = np.load(ROOT / "tests" / "golden" / "oracle_n_160x120.npz")
    ok, T, ex = O.align_depth_pair(g["frames"][1], g["frames"][0], tuple(g["intr"]))
    assert ok and np.array_equal(T, GOLD["depth_pair_T"])
    assert [ex["n_src"], ex["n_dst"]] == GOLD["depth_pair_n"].tolist()
    et, er = synth.pose_error(T, g["gt"][0])
    assert et < 0.05 and er < 0.05          # point-to-point on 5 cm voxels: centimetre-level, as the reference
    oks, Ts = O.align_depth_pairs(g["frames"][1:3], g["frames"][0:2], tuple(g["intr"]), n_threads=2)
    assert oks.all() and np.array_equal(Ts[0], T)


def test_backprojection_maps_invalid_depth_to_the_origin():
    d = np.array([[0, 1000], [2000, 0]], dtype=np.uint16)
    c = O.backproject(d, (100.0, 100.0, 0.5, 0.5))
    assert np.array_equal(c[0], [0, 0, 0]) and np.array_equal(c[3], [0, 0, 0])   # rs_driver.cpp:83-88
    assert np.allclose(c[1], [0.005, -0.005, 1.0]) and np.allclose(c[2], [-0.01, 0.01, 2.0])
