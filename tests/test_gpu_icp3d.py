"""GPU: the cloud-based engine (rst_icp3d_pairs) against Oracle-R, the restatement of the reference's
AlignIcp3d (align_icp.cpp:73-167). NN indices and weights bit-exact for one iteration from the same
pose, cross-covariance within 1e-4 relative, final poses within 1e-4 m / 1e-4 rad."""
import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O
from realsensetracker_b200 import Aligner, AlignIcp3d, synth

pytestmark = pytest.mark.gpu
GOLD = np.load(ROOT / "tests" / "golden" / "oracle_r.npz")
GN = np.load(ROOT / "tests" / "golden" / "oracle_n_160x120.npz")


@pytest.fixture(scope="module")
def al():
    a = Aligner(16, 16, 2, 1)
    yield a
    a.close()


def depth_clouds(i_src, i_dst, voxel=0.05):
    intr = tuple(GN["intr"])
    s = O.downsample_voxel(O.remove_nans(O.backproject(GN["frames"][i_src], intr)), voxel)
    d = O.downsample_voxel(O.remove_nans(O.backproject(GN["frames"][i_dst], intr)), voxel)
    return s, d


@pytest.mark.parametrize("case", ["random", "depth"])
def test_single_iteration_is_bit_exact(al, case):
    if case == "random":
        src, dst, T0 = GOLD["src"], GOLD["dst"], np.eye(4)
    else:
        src, dst = depth_clouds(1, 0)
        T0 = synth.make_pose(synth.rotvec_to_R([0.01, -0.02, 0.01]), [0.02, 0.01, -0.03])
    ok_o, T_o, ex_o = O.align_icp3d(src, dst, 1, T0=T0, details=True)
    ok_g, T_g, ex_g = al.icp3d_pairs([src], [dst], 1, T0=T0, details=True)
    assert ok_g[0] == ok_o
    assert np.array_equal(ex_g[0]["nbrs"], ex_o["nbrs"]), f"{(ex_g[0]['nbrs'] != ex_o['nbrs']).sum()} NN indices differ"
    assert np.array_equal(ex_g[0]["weights"], ex_o["weights"])
    assert np.allclose(ex_g[0]["cov"], ex_o["cov"], rtol=1e-4, atol=1e-4 * np.abs(ex_o["cov"]).max())
    assert abs(ex_g[0]["mean_cost"] - ex_o["mean_cost"]) <= 1e-5 * max(ex_o["mean_cost"], 1e-6)
    dt, dr = synth.pose_error(T_g[0], T_o)
    assert dt < 1e-5 and dr < 1e-5


def test_full_run_matches_oracle_r_and_known_motion(al):
    src, dst = GOLD["src"], GOLD["dst"]
    ok_o, T_o, ex_o = O.align_icp3d(src, dst, 128, details=True)
    ok_g, T_g, ex_g = al.icp3d_pairs([src], [dst], 128, details=True)
    dt, dr = synth.pose_error(T_g[0], T_o)
    assert ok_g[0] and dt < 1e-4 and dr < 1e-4, (dt, dr)
    assert synth.pose_error(T_g[0], GOLD["T_small"]) < (1e-4, 1e-4)       # rs_align_app.cpp:257-263 self-test
    assert (ex_g[0]["nbrs"] == ex_o["nbrs"]).mean() > 0.999
    assert abs(ex_g[0]["mu"] - 1.0 / 1.4 ** 15) < 1e-6
    ok1, T1 = AlignIcp3d(src, dst, 128)                                     # the reference's signature
    assert ok1 and np.array_equal(T1, T_g[0])


def test_batch_of_depth_clouds_and_determinism(al):
    pairs = [depth_clouds(1, 0), depth_clouds(2, 1), depth_clouds(2, 0)]
    S, D = [p[0] for p in pairs], [p[1] for p in pairs]
    ok, T = al.icp3d_pairs(S, D, 128)
    ok2, T2 = al.icp3d_pairs(S, D, 128)
    assert ok.all() and np.array_equal(T, T2)
    for i in range(3):
        ok_o, T_o = O.align_icp3d(S[i], D[i], 128)
        dt, dr = synth.pose_error(T[i], T_o)
        assert dt < 1e-4 and dr < 1e-4, (i, dt, dr)
    one_ok, one_T = al.icp3d_pairs(S[1:2], D[1:2], 128)
    assert np.array_equal(one_T[0], T[1])                                   # independent of the batch


def test_failure_and_edge_cases(al):
    src, dst = GOLD["src"], GOLD["dst"]
    ok, T = al.icp3d_pairs([src[:2]], [dst], 5)
    assert not ok[0] and np.array_equal(T[0], np.eye(4))                    # < 3 points -> false, pose untouched
    ok, T = al.icp3d_pairs([src], [dst], 0, T0=GOLD["T_small"])
    assert np.allclose(T[0], GOLD["T_small"], atol=1e-7)                    # 0 iterations: pose passes through
    far = src + np.float32(50.0)                                            # every query far outside the dst grid
    ok_o, T_o, ex_o = O.align_icp3d(far, dst, 1, details=True)
    ok_g, T_g, ex_g = al.icp3d_pairs([far], [dst], 1, details=True)
    assert np.array_equal(ex_g[0]["nbrs"], ex_o["nbrs"])
    dup = np.concatenate([dst, dst[:50]])                                   # exact ties -> lowest index, like the oracle
    ok_o, T_o, ex_o = O.align_icp3d(src, dup, 1, details=True)
    ok_g, T_g, ex_g = al.icp3d_pairs([src], [dup], 1, details=True)
    assert np.array_equal(ex_g[0]["nbrs"], ex_o["nbrs"])


def test_depth_to_cloud_is_bit_exact_and_depth_icp_matches_the_reference_pipeline(al):
    """rst_icp3d_depth = the caller's sequence rs_replay_app.cpp:229,246-251 on the device."""
    frames, intr = GN["frames"], tuple(GN["intr"])
    src_idx, dst_idx = [1, 2], [0, 1]                      # frame-to-frame sequence
    ok, T, mean_cost, counts = al.icp3d_depth(frames, src_idx, dst_idx, intr, voxel=0.05, max_iter=128)
    for f in range(3):
        want = O.downsample_voxel(O.remove_nans(O.backproject(frames[f], intr)), 0.05)
        assert counts[f] == len(want)
        got = al.icp3d_read_cloud(f, int(counts[f]))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"cloud of frame {f} differs"
    for i in range(2):
        ok_o, T_o, ex = O.align_depth_pair(frames[src_idx[i]], frames[dst_idx[i]], intr)
        dt, dr = synth.pose_error(T[i], T_o)
        assert ok[i] == ok_o and dt < 1e-4 and dr < 1e-4, (i, dt, dr)
        assert abs(mean_cost[i] - ex["mean_cost"]) < 1e-5
    # no decimation: the full-resolution cloud, invalid pixels at the origin (rs_driver.cpp:83-88)
    z = frames[0].copy(); z[:40] = 0
    ok, T, mc, counts = al.icp3d_depth(np.stack([z, frames[1]]), [1], [0], intr, voxel=0.0, max_iter=2)
    assert counts.tolist() == [160 * 120, 160 * 120]
    cloud = al.icp3d_read_cloud(0, 160 * 120)
    assert np.array_equal(cloud, O.backproject(z, intr)) and (cloud[:40 * 160] == 0).all()


@pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref/libref.so not present")
def test_gpu_against_the_compiled_reference_source(al):
    """The CUDA cloud engine against the reference's own align_icp.cpp (compiled with stand-in headers)."""
    src, dst = GOLD["src"], GOLD["dst"]
    for iters in (1, 128):
        ok_r, T_r = O.ref_align_icp3d(src, dst, iters)
        ok_g, T_g = al.icp3d_pairs([src], [dst], iters)
        dt, dr = synth.pose_error(T_g[0], T_r)
        assert ok_g[0] == ok_r and dt < 1e-4 and dr < 1e-4, (iters, dt, dr)


def test_solve_kabsch_on_the_device(al):
    """rst_solve_kabsch vs Oracle-R (and, through it, the compiled reference): align_icp.cpp:18-71."""
    src, dst = GOLD["src"], GOLD["dst_big"]
    rng = np.random.default_rng(4)
    pairs = np.stack([np.arange(len(src)), np.arange(len(src))], 1)[rng.permutation(len(src))[:400]]
    for w in (None, GOLD["kabsch_w"][:400]):
        ok_o, T_o = O.solve_kabsch(src, dst, pairs, w)
        ok_g, T_g = al.solve_kabsch(src, dst, pairs, w)
        dt, dr = synth.pose_error(T_g, T_o)
        assert ok_g and ok_o and dt < 1e-5 and dr < 1e-5, (dt, dr)
        assert synth.pose_error(T_g, GOLD["T_big"]) < (1e-4, 1e-4)          # the known rotation of rs_align_app.cpp:257-263
    ok_g, _ = al.solve_kabsch(src[:2], dst, pairs[:1] * 0)
    assert ok_g is False                                                     # < 3 points (:23-25)
    # the device SVD (division-free one-sided Jacobi) against LAPACK, not against an oracle that shares its algorithm:
    # R = U V^T of the fp64 cross-covariance by numpy.linalg.svd, for well- and ill-conditioned (nearly planar) sets
    for flat in (1.0, 1e-3):
        s2 = (rng.standard_normal((500, 3)) * [1.0, 0.7, flat]).astype(np.float32)
        R_true = synth.rotvec_to_R([0.4, -0.9, 0.3])
        d2 = (s2 @ R_true.T.astype(np.float32) + np.float32([0.3, -0.2, 1.1])).astype(np.float32)
        pr = np.stack([np.arange(500), np.arange(500)], 1)
        ok_g, T_g = al.solve_kabsch(s2, d2, pr)
        sm, dm = s2.mean(0, dtype=np.float64).astype(np.float32), d2.mean(0, dtype=np.float64).astype(np.float32)
        cov = ((d2 - dm)[:, :, None] * (s2 - sm)[:, None, :]).astype(np.float32).astype(np.float64).sum(0)
        U, _, Vt = np.linalg.svd(cov)
        R_np = U @ Vt
        assert ok_g and np.abs(T_g[:3, :3] - R_np).max() < 2e-6, (flat, np.abs(T_g[:3, :3] - R_np).max())
        assert np.abs(T_g[:3, :3] - R_true).max() < (2e-6 if flat == 1.0 else 2e-3)
    # Kabsch initialiser -> AlignIcp3d, the order rs_align_app.cpp:295-303 uses
    ok_k, T_k = al.solve_kabsch(src, dst, pairs)
    ok_i, T_i = al.icp3d_pairs([src], [dst], 16, T0=T_k)
    assert ok_i[0] and synth.pose_error(T_i[0], GOLD["T_big"]) < (1e-4, 1e-4)


def test_cloud_normals_on_the_device(al):
    """rst_cloud_normals vs the reference's own ComputeNormals + OrientNormals (compiled, oracle/_ref) when
    present, and vs an analytic plane always (point_cloud_utils.cpp:176-216)."""
    rng = np.random.default_rng(0)
    xy = rng.uniform(-1, 1, size=(3000, 2))
    n = np.array([0.3, -0.2, -1.0]); n /= np.linalg.norm(n)
    z = (-2.0 - xy @ n[:2]) / n[2]
    plane = np.column_stack([xy, z]).astype(np.float32)
    got = al.cloud_normals(plane, k=16)
    assert np.allclose(np.abs(got @ n), 1.0, atol=1e-3)
    assert ((got * plane).sum(1) <= 0).all()                               # orientation rule, viewpoint = origin
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-4)
    flipped = al.cloud_normals(plane, k=16, viewpoint=(0.0, 0.0, 10.0))   # seen from behind: normals flip
    assert ((flipped * (plane - np.float32([0, 0, 10]))).sum(1) <= 0).all() and np.allclose(flipped, -got, atol=1e-4)
    # a depth-derived cloud (curved + planar parts) against the compiled reference
    if O.ref_lib() is not None:
        cloud = depth_clouds(1, 0)[0][:4000]
        ref = O.ref_normals(cloud, k=16)
        gpu = al.cloud_normals(cloud, k=16)
        cosang = np.abs((ref * gpu).sum(1))
        assert np.median(cosang) > 0.99999 and (cosang > 0.999).mean() > 0.97   # near-isotropic neighbourhoods may differ
        assert ((gpu * cloud).sum(1) <= 1e-6).all()
        assert ((ref * gpu).sum(1) > 0).mean() > 0.97                      # same orientation


def test_cluster_per_pair_agrees_with_one_cta_per_pair(al):
    """Small batches run a thread-block cluster per pair (the source points split over its CTAs, sums through
    distributed shared memory). Same neighbours bit for bit after one iteration from the same pose; the 128-iteration
    pose within fp64 partial-sum round-off of the one-CTA result; a two-pair batch, ragged cloud sizes and a
    degenerate pair (fewer than 3 points) take the same path."""
    src, dst = depth_clouds(1, 0)
    src2, dst2 = GOLD["src"], GOLD["dst"]
    try:
        al.set_icp3d_cluster(1)
        ok1, T1, ex1 = al.icp3d_pairs([src, src2], [dst, dst2], 1, details=True)
        okf, Tf, exf = al.icp3d_pairs([src, src2], [dst, dst2], 128, details=True)
        for c in (2, 8, 16, 0):
            al.set_icp3d_cluster(c)
            ok_c, T_c, ex_c = al.icp3d_pairs([src, src2], [dst, dst2], 1, details=True)
            for i in range(2):
                assert np.array_equal(ex_c[i]["nbrs"], ex1[i]["nbrs"]) and np.array_equal(ex_c[i]["weights"], ex1[i]["weights"])
                assert np.allclose(ex_c[i]["cov"], ex1[i]["cov"], rtol=1e-12, atol=1e-12 * np.abs(ex1[i]["cov"]).max())
            ok_c, T_c, ex_c = al.icp3d_pairs([src, src2], [dst, dst2], 128, details=True)
            for i in range(2):
                dt, dr = synth.pose_error(T_c[i], Tf[i])
                assert ok_c[i] == okf[i] and dt < 1e-5 and dr < 1e-5, (c, i, dt, dr)
                assert abs(ex_c[i]["mean_cost"] - exf[i]["mean_cost"]) <= 1e-5 * max(exf[i]["mean_cost"], 1e-6)
        al.set_icp3d_cluster(16)
        ok_d, T_d = al.icp3d_pairs([src[:2]], [dst], 8, T0=np.eye(4))   # fewer than 3 points: false, pose untouched
        assert not ok_d[0] and np.array_equal(T_d[0], np.eye(4, dtype=np.float32))
    finally:
        al.set_icp3d_cluster(0)


def test_neighbour_cache_changes_nothing_but_time(al):
    """rst_set_icp3d_cache: the cached neighbour is kept only when the triangle inequality proves a search would return
    it, so 128 iterations with the cache (any margin setting, any cluster size) give the same neighbours, weights,
    covariance and pose as 128 iterations that search every point every time — bit for bit."""
    src, dst = depth_clouds(1, 0)
    src2, dst2 = GOLD["src"], GOLD["dst"]
    T0 = synth.make_pose(synth.rotvec_to_R([0.01, -0.02, 0.01]), [0.02, 0.01, -0.03])
    try:
        for cl in (1, 4):
            al.set_icp3d_cluster(cl)
            al.set_icp3d_cache(0.0, 0.0, 0.0)
            ok0, Ta, ex0 = al.icp3d_pairs([src, src2], [dst, dst2], 128, T0=T0, details=True)
            for setting in ((1.0, 0.05, 0.2), (4.0, 0.05, 0.5), (0.0, 0.0, 0.01), (8.0, 0.5, 2.0)):
                al.set_icp3d_cache(*setting)
                for cell in (0.1, 0.0):
                    ok1, Tb, ex1 = al.icp3d_pairs([src, src2], [dst, dst2], 128, T0=T0, grid_cell=cell, details=True)
                    assert np.array_equal(Ta, Tb), (cl, setting, cell)
                    for i in range(2):
                        assert np.array_equal(ex0[i]["nbrs"], ex1[i]["nbrs"]) and np.array_equal(ex0[i]["weights"], ex1[i]["weights"])
                        assert np.array_equal(ex0[i]["cov"], ex1[i]["cov"]) and ex0[i]["mean_cost"] == ex1[i]["mean_cost"]
        with pytest.raises(Exception):
            al.set_icp3d_cache(1.0, 2.0, 1.0)       # lo > hi
    finally:
        al.set_icp3d_cluster(0)
        al.set_icp3d_cache()


def test_neighbour_cache_ties_long_runs_and_large_motion(al):
    """The cases the cache has to get right by NOT skipping: duplicated target points (ties go to the lowest index, the
    runner-up distance equals the best), 300 iterations (the iteration tag of a cache entry wraps at 256; entries older
    than 250 iterations search again), an initial misalignment of 20 cm / 12 degrees (every early iteration moves every
    point further than its proven radius), a ragged source size — each bit-identical to searching every time."""
    rng = np.random.default_rng(7)
    src, dst = depth_clouds(1, 0)
    dst_dup = np.concatenate([dst, dst[::3], dst[::7]]).astype(np.float32)           # exact duplicates, higher indices
    T_far = synth.make_pose(synth.rotvec_to_R([0.12, -0.15, 0.08]), [0.2, -0.1, 0.15])
    cases = [(src[:4001], dst_dup, np.eye(4), 64), (src, dst, T_far, 128), (src[:1500], dst, np.eye(4), 300),
             (src[:5], dst, np.eye(4), 10), (src[:1025], dst[:50], np.eye(4), 20),      # CTAs of a cluster without a point; tiny target

             ((rng.random((3000, 3)) * 2 - 1).astype(np.float32), (rng.random((2500, 3)) * 2 - 1).astype(np.float32), np.eye(4), 40)]
    try:
        for cl in (1, 8):
            al.set_icp3d_cluster(cl)
            for s_, d_, T0, it in cases:
                al.set_icp3d_cache(0.0, 0.0, 0.0)
                ok0, Ta, ex0 = al.icp3d_pairs([s_], [d_], it, T0=T0, details=True)
                searched0, queried0 = al.icp3d_cache_stats()
                run0, asked0 = al.icp3d_iteration_stats()
                assert queried0 == len(s_) * it and asked0 == it and searched0 == len(s_) * run0   # cache off: every iteration that ran searched every point
                al.set_icp3d_cache()
                ok1, Tb, ex1 = al.icp3d_pairs([s_], [d_], it, T0=T0, details=True)
                searched1, queried1 = al.icp3d_cache_stats()
                assert queried1 == queried0 and len(s_) <= searched1 <= queried1
                assert np.array_equal(Ta, Tb) and ok0[0] == ok1[0], (cl, it)
                assert np.array_equal(ex0[0]["nbrs"], ex1[0]["nbrs"]) and np.array_equal(ex0[0]["weights"], ex1[0]["weights"])
                assert np.array_equal(ex0[0]["cov"], ex1[0]["cov"]) and ex0[0]["mean_cost"] == ex1[0]["mean_cost"]
        # duplicates: the reported neighbour is the lowest index among the tied points
        ok, T, ex = al.icp3d_pairs([src[:4001]], [dst_dup], 3, details=True)
        first_of = {tuple(p): i for i, p in reversed(list(enumerate(map(tuple, dst_dup))))}
        assert all(first_of[tuple(dst_dup[j])] == j for j in ex[0]["nbrs"])
    finally:
        al.set_icp3d_cluster(0)
        al.set_icp3d_cache()


def test_fixed_point_skip_changes_nothing_but_time(al):
    """rst_set_icp3d_fixed_point_skip: once an iteration returns the pose it was given bit for bit (and its SVD warm-start
    basis unchanged), the iterations up to the next change of mu would repeat it exactly and are not run. Same
    neighbours, weights, covariance, cost, mu and pose as running every iteration — for odd iteration counts, counts
    that end inside a skipped stretch, one CTA per pair and clusters, cache on and off; and iterations are saved."""
    src, dst = depth_clouds(1, 0)
    src2, dst2 = GOLD["src"], GOLD["dst"]
    T0 = synth.make_pose(synth.rotvec_to_R([0.01, -0.02, 0.01]), [0.02, 0.01, -0.03])
    saved = 0
    try:
        for cl in (1, 4):
            al.set_icp3d_cluster(cl)
            for cache in ((1.0, 0.05, 0.2), (0.0, 0.0, 0.0)):
                al.set_icp3d_cache(*cache)
                for it in (128, 61, 30, 23, 9):
                    al.set_icp3d_fixed_point_skip(False)
                    ok0, Ta, ex0 = al.icp3d_pairs([src, src2], [dst, dst2], it, T0=T0, details=True)
                    assert al.icp3d_iteration_stats() == (2 * it, 2 * it)
                    al.set_icp3d_fixed_point_skip(True)
                    ok1, Tb, ex1 = al.icp3d_pairs([src, src2], [dst, dst2], it, T0=T0, details=True)
                    run, asked = al.icp3d_iteration_stats()
                    assert asked == 2 * it and 2 <= run <= asked
                    saved += asked - run
                    assert np.array_equal(Ta, Tb) and np.array_equal(ok0, ok1), (cl, cache, it)
                    for i in range(2):
                        assert np.array_equal(ex0[i]["nbrs"], ex1[i]["nbrs"]) and np.array_equal(ex0[i]["weights"], ex1[i]["weights"])
                        assert np.array_equal(ex0[i]["cov"], ex1[i]["cov"]) and ex0[i]["mean_cost"] == ex1[i]["mean_cost"]
                        assert ex0[i]["mu"] == ex1[i]["mu"]
        assert saved > 0, "no iteration was ever skipped on converging pairs"
    finally:
        al.set_icp3d_cluster(0)
        al.set_icp3d_cache()
        al.set_icp3d_fixed_point_skip(True)
