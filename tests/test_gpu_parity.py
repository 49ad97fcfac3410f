"""GPU parity: the CUDA path (through the C ABI) against Oracle-N on the same seeded inputs.

Bars (BASELINE.json north_star): pyramid / geometry map / association indices bit-exact;
normal equations within 1e-4 relative; poses within 1e-4 m and 1e-4 rad.
"""
import numpy as np
import pytest

from oracle import oracle as O
from realsensetracker_b200 import Aligner, default_params, synth
from realsensetracker_b200 import _native as N

pytestmark = pytest.mark.gpu

TOL_M = 1e-4    # metres
TOL_RAD = 1e-4  # radians
TOL_NE = 1e-4   # relative, normal equations


def oparams(**kw):
    return O.default_params(**kw)


def both_params(**kw):
    return default_params(**kw), O.default_params(**kw)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def jtr_error(st_g, st_o):
    """max_k |b_g[k] - b_o[k]| / sqrt(A_kk * sum w r^2): b = J^T r is a sum of products J_k r whose total magnitude is
    bounded by sqrt(sum J_k^2 * sum r^2) = sqrt(A_kk * sum_wr2) (Cauchy-Schwarz). Near the solution b itself cancels
    to ~0, so an error relative to max|b| is meaningless there; relative to the size of its summands the 1e-4
    contract of the normal equations is well defined at every pose."""
    A = np.array(st_o.A[:]); bo = np.array(st_o.b[:]); bg = np.array(st_g.b[:])
    diag = A[[0, 6, 11, 15, 18, 20]]
    scale = np.sqrt(np.maximum(diag * st_o.sum_wr2, 1e-300))
    return float(np.max(np.abs(bg - bo) / scale))


def oracle_levels(frame, intr, P):
    lv, d = [], frame
    for l in range(P.num_levels):
        L = O.level_info(intr, frame.shape[1], frame.shape[0], l)
        if l > 0:
            d = O.pyr_down(d, P.pyr_depth_tol)
        lv.append((L, d, O.geometry(d, L, P)))
    return lv


@pytest.fixture(scope="module", params=["640", "small"])
def staged(request, seq640, seq_small):
    frames, gt, intr = seq640 if request.param == "640" else seq_small
    n, h, w = frames.shape
    P, Po = both_params()
    al = Aligner(w, h, n, n)
    al.begin(w, h, intr, P)
    al.upload(frames)
    al.preprocess(0, n)
    yield al, frames, gt, intr, P, Po
    al.close()


def test_pyramid_bit_exact(staged):
    al, frames, gt, intr, P, Po = staged
    for slot in range(2):
        d = frames[slot]
        assert np.array_equal(al.read_depth(slot, 0), d)
        for l in range(1, P.num_levels):
            d = O.pyr_down(d, Po.pyr_depth_tol)
            got = al.read_depth(slot, l)
            assert got.shape == d.shape
            assert np.array_equal(got, d), f"pyramid level {l} differs"


def test_level_intrinsics_match(staged):
    al, frames, gt, intr, P, Po = staged
    for l in range(P.num_levels):
        w, h, pitch, K = al.level_info(l)
        L = O.level_info(intr, frames.shape[2], frames.shape[1], l)
        assert (w, h) == (L.w, L.h)
        assert np.array_equal(np.float32(K), np.float32([L.fx, L.fy, L.cx, L.cy]))
        assert pitch % 8 == 0 and pitch >= w


def test_geometry_map_bit_exact(staged):
    al, frames, gt, intr, P, Po = staged
    for slot in range(2):
        for l, (L, d, G) in enumerate(oracle_levels(frames[slot], intr, Po)):
            got = al.read_geometry(slot, l)
            assert got.shape == G.shape
            valid = G[..., 3] > 0
            assert valid.mean() > 0.3
            assert np.array_equal(got.view(np.uint32), G.view(np.uint32)), f"geometry level {l} differs"


@pytest.mark.parametrize("pose_kind", ["identity", "gt", "perturbed"])
def test_association_bit_exact_and_normal_equations(staged, pose_kind):
    al, frames, gt, intr, P, Po = staged
    T = {"identity": np.eye(4), "gt": gt[0],
         "perturbed": gt[0] @ synth.make_pose(synth.rotvec_to_R([0.01, -0.02, 0.015]), [0.02, -0.01, 0.03])}[pose_kind]
    T = np.asarray(T, dtype=np.float32).astype(np.float64)
    src_lv = oracle_levels(frames[1], intr, Po)
    dst_lv = oracle_levels(frames[0], intr, Po)
    for l in range(P.num_levels):
        L, ds, _ = src_lv[l]
        idx_o, st_o = O.evaluate(ds, None, dst_lv[l][2], L, Po, T)
        idx_g, st_g = al.evaluate(1, 0, l, T)
        assert np.array_equal(idx_g, idx_o), f"association differs at level {l}: {(idx_g != idx_o).sum()} px"
        assert st_g.count == st_o.count == int((idx_o >= 0).sum())
        assert st_o.count > 500
        assert rel_err(st_g.A[:], st_o.A[:]) < TOL_NE
        e_b = jtr_error(st_g, st_o)
        print(f"level {l} {pose_kind}: J^T r error / sqrt(A_kk sum_wr2) = {e_b:.2e}, relative to max|b| = {rel_err(st_g.b[:], st_o.b[:]):.2e}")
        assert e_b < TOL_NE
        assert abs(st_g.sum_wr2 - st_o.sum_wr2) <= TOL_NE * st_o.sum_wr2 + 1e-12


@pytest.mark.parametrize("kind,scale", [(N.RST_ROBUST_HUBER, 0.002), (N.RST_ROBUST_GEMAN_MCCLURE, 1e-4)])
def test_robust_weights_and_normal_gate(seq_small, kind, scale):
    frames, gt, intr = seq_small
    n, h, w = frames.shape
    kw = dict(robust_kind=kind, robust_scale=scale, normal_cos_min=0.9)
    P, Po = both_params(**kw)
    al = Aligner(w, h, n, n)
    try:
        al.begin(w, h, intr, P)
        al.upload(frames)
        al.preprocess(0, n)
        T = np.eye(4)
        src_lv, dst_lv = oracle_levels(frames[2], intr, Po), oracle_levels(frames[1], intr, Po)
        for l in range(P.num_levels):
            L, ds, Gs = src_lv[l]
            idx_o, st_o = O.evaluate(ds, Gs, dst_lv[l][2], L, Po, T)
            idx_g, st_g = al.evaluate(2, 1, l, T)
            assert np.array_equal(idx_g, idx_o)
            assert rel_err(st_g.A[:], st_o.A[:]) < TOL_NE
            assert jtr_error(st_g, st_o) < TOL_NE
            assert abs(st_g.sum_wr2 - st_o.sum_wr2) <= TOL_NE * st_o.sum_wr2
    finally:
        al.close()


def _check_pose(Tg, To, gt=None, gt_tol=(2e-3, 2e-3)):
    dt, dr = synth.pose_error(Tg, To)
    assert dt < TOL_M and dr < TOL_RAD, f"GPU vs oracle pose: {dt} m, {dr} rad"
    if gt is not None:
        et, er = synth.pose_error(Tg, gt)
        assert et < gt_tol[0] and er < gt_tol[1], f"GPU vs ground truth: {et} m, {er} rad"


def test_pose_parity_pairs_640(seq640):
    frames, gt, intr = seq640
    P, Po = both_params()
    al = Aligner(640, 480, 8, 4)
    try:
        Tg, st = al.align_pairs(frames[1:4], frames[0:3], intr, P)
        for i in range(3):
            To, so = O.align_pair(frames[i + 1], frames[i], intr, Po)
            assert st[i].status == 0 and so.status == 0
            assert st[i].iterations == so.iterations == 19
            _check_pose(Tg[i], To, gt[i])
            assert st[i].count == so.count or abs(st[i].count - so.count) <= 4
            assert rel_err(st[i].A[:], so.A[:]) < TOL_NE
            assert abs(st[i].rmse - so.rmse) < 1e-6
    finally:
        al.close()


def test_sequence_equals_pairs_and_is_deterministic(seq640):
    frames, gt, intr = seq640
    P, _ = both_params()
    al = Aligner(640, 480, 10, 5)
    try:
        Ts, _ = al.align_sequence(frames, intr, P)
        Ts2, _ = al.align_sequence(frames, intr, P)
        Tp, _ = al.align_pairs(frames[1:], frames[:-1], intr, P)
        T1, _ = al.align_pairs(frames[2:3], frames[1:2], intr, P)   # batch of one
        assert np.array_equal(Ts, Ts2), "two runs differ bitwise"
        assert np.array_equal(Ts, Tp), "sequence API and pairs API differ bitwise"
        assert np.array_equal(T1[0], Ts[1]), "result depends on the batch size"
    finally:
        al.close()


def test_initial_pose_is_used_and_large_motion(seq640):
    """KAT from the reference's disabled self-test (rs_align_app.cpp:257-263): known rotation
    R_x(0.1) R_y(-0.2) R_z(0.25), scaled to a trackable magnitude, recovered from an initial guess."""
    frames, gt, intr = seq640
    scene = synth.Scene(0)
    Rk = synth.rot_xyz(0.1 * 0.3, -0.2 * 0.3, 0.25 * 0.3)
    T_src = synth.make_pose(Rk, [0.03, -0.02, 0.04])
    src = scene.render(T_src, 640, 480)
    dst = scene.render(np.eye(4), 640, 480)
    T_gt = synth.relative_pose(np.eye(4), T_src)
    kw = dict(iters=[10, 10, 10], dist_max=0.5)
    P, Po = both_params(**kw)
    T0 = T_gt @ synth.make_pose(synth.rotvec_to_R([0.01, 0.01, -0.01]), [0.01, -0.01, 0.01])
    al = Aligner(640, 480, 2, 1)
    try:
        for init in (None, T0):
            Tg, st = al.align_pairs(src[None], dst[None], intr, P, T0=init)
            To, so = O.align_pair(src, dst, intr, Po, T0=init)
            assert st[0].status == 0
            _check_pose(Tg[0], To, T_gt, gt_tol=(3e-3, 2e-3))
    finally:
        al.close()


def test_invalid_depth_and_failure_status(seq_small):
    frames, gt, intr = seq_small
    n, h, w = frames.shape
    P, Po = both_params()
    al = Aligner(w, h, 4, 2)
    try:
        # 30 % invalid (Bernoulli + blocks) still aligns and matches the oracle
        rng = np.random.default_rng(1)
        src, dst = frames[1].copy(), frames[0].copy()
        for f in (src, dst):
            f[rng.random(f.shape) < 0.2] = 0
            for _ in range(12):
                y, x = rng.integers(0, h - 16), rng.integers(0, w - 16)
                f[y:y + 16, x:x + 16] = 0
        Tg, st = al.align_pairs(src[None], dst[None], intr, P)
        To, so = O.align_pair(src, dst, intr, Po)
        assert st[0].status == 0
        _check_pose(Tg[0], To, gt[0], gt_tol=(5e-3, 5e-3))
        # all-invalid source: too few associations -> failure status, pose unchanged (align_icp.cpp:77-79)
        zero = np.zeros_like(src)
        Tg, st = al.align_pairs(zero[None], dst[None], intr, P)
        To, so = O.align_pair(zero, dst, intr, Po)
        assert st[0].status == so.status == N.RST_STATUS_TOO_FEW
        assert np.array_equal(Tg[0], np.eye(4))
        assert st[0].count == 0
    finally:
        al.close()


def test_device_resident_frames_match_host_path(seq640):
    import torch
    frames, gt, intr = seq640
    P, _ = both_params()
    n = frames.shape[0]
    al = Aligner(640, 480, n, n)
    try:
        Th, _ = al.align_sequence(frames, intr, P)
        d = torch.from_numpy(frames.astype(np.int16)).cuda()
        al.begin(640, 480, intr, P)
        al.set_frames_device(d.data_ptr(), n, 640, 640 * 480)
        al.preprocess(0, n)
        Td, st = al.align_slots(np.arange(1, n), np.arange(0, n - 1))
        assert np.array_equal(Th, Td)
        with pytest.raises(Exception):
            al.set_frames_device(d.data_ptr() + 2, n, 640, 640 * 480)
    finally:
        al.close()


def test_capacity_and_argument_errors():
    al = Aligner(64, 48, 2, 1)
    try:
        big = np.zeros((1, 96, 128), dtype=np.uint16)
        with pytest.raises(Exception) as e:
            al.align_pairs(big, big, (100, 100, 64, 48))
        assert "CAPACITY" in str(e.value)
        ok = np.zeros((3, 48, 64), dtype=np.uint16)
        with pytest.raises(Exception) as e:
            al.align_pairs(ok, ok, (50, 50, 32, 24))
        assert "CAPACITY" in str(e.value)
        with pytest.raises(Exception) as e:
            al.align_pairs(ok[:1], ok[:1], (50, 50, 32, 24), default_params(num_levels=9))
        assert "INVALID_ARG" in str(e.value)
    finally:
        al.close()


def test_async_submit_wait_equals_blocking_call(seq640):
    """Two contexts driven alternately (the streaming form): same bits as the blocking call."""
    frames, gt, intr = seq640
    P, _ = both_params()
    a, b = Aligner(640, 480, 8, 4), Aligner(640, 480, 5, 4)
    try:
        Tref, sref = a.align_sequence(frames, intr, P)
        a.submit_sequence(frames, intr, P)
        b.submit_sequence(frames[::-1].copy(), intr, P)
        Ta, sa = a.wait()
        Tb, sb = b.wait()
        assert np.array_equal(Ta, Tref) and [s.count for s in sa] == [s.count for s in sref]
        Trev, _ = a.align_sequence(frames[::-1].copy(), intr, P)
        assert np.array_equal(Tb, Trev)
        for chunk in (2, 3):      # chunked two-stream pipeline inside one call: same bits
            a.set_pipeline_chunk(chunk)
            Tc, _ = a.align_sequence(frames, intr, P)
            assert np.array_equal(Tc, Tref)
            Tp, _ = a.align_pairs(frames[1:], frames[:-1], intr, P)
            assert np.array_equal(Tp, Tref)
    finally:
        a.close(); b.close()


def test_latency_tiling_matches_oracle_and_is_batch_invariant(seq640):
    frames, gt, intr = seq640
    P = default_params(tiling=N.RST_TILING_LATENCY)
    Po = O.default_params()
    al = Aligner(640, 480, 8, 4)
    try:
        T, st = al.align_pairs(frames[1:4], frames[0:3], intr, P)
        T1, _ = al.align_pairs(frames[2:3], frames[1:2], intr, P)
        assert np.array_equal(T1[0], T[1])
        for i in range(3):
            To, so = O.align_pair(frames[i + 1], frames[i], intr, Po)
            _check_pose(T[i], To, gt[i])
        with pytest.raises(Exception):
            al.align_pairs(frames[1:2], frames[0:1], intr, default_params(tiling=7))
    finally:
        al.close()


def test_abi_edge_cases_on_the_device(seq_small):
    """Empty batches, short sequences, non-finite priors: defined behaviour, no crash."""
    import ctypes as C
    frames, gt, intr = seq_small
    n, h, w = frames.shape
    P = default_params()
    al = Aligner(w, h, 4, 2)
    try:
        lib, ctx = al._lib, al._ctx
        K = N.Intrinsics(*intr)
        from realsensetracker_b200.align import _frames
        poses = np.tile(np.eye(4, dtype=np.float32).reshape(16), (2, 1))
        assert lib.rst_align_pairs(ctx, _frames(frames[:1]), _frames(frames[:1]), 0, C.byref(K), C.byref(P), poses.ctypes.data, None) == N.RST_OK
        assert lib.rst_align_sequence(ctx, _frames(frames[:1]), 1, C.byref(K), C.byref(P), poses.ctypes.data, None) == N.RST_OK
        assert lib.rst_align_pairs(ctx, None, None, 1, C.byref(K), C.byref(P), poses.ctypes.data, None) == N.RST_ERR_INVALID_ARG
        assert b"null" in lib.rst_last_error(ctx)
        # a NaN prior: the pair reports NON_FINITE / TOO_FEW instead of poisoning its neighbour in the batch
        T0 = np.stack([np.eye(4), np.eye(4)])
        T0[0, 0, 3] = np.nan
        T, st = al.align_pairs(frames[1:3], frames[0:2], intr, P, T0=T0)
        assert st[0].status != 0
        assert st[1].status == 0 and synth.pose_error(T[1], gt[1])[0] < 5e-3
        # stats are optional
        assert lib.rst_align_pairs(ctx, _frames(frames[1:2]), _frames(frames[0:1]), 1, C.byref(K), C.byref(P), poses.ctypes.data, None) == N.RST_OK
        assert al.launch_count > 0
    finally:
        al.close()


def test_convergence_early_exit_matches_oracle(seq640):
    """converge_eps > 0: a pair leaves a level once an update is below the threshold (the reference has no
    such test, align_icp.cpp:92 — fixed count stays the default)."""
    frames, gt, intr = seq640
    kw = dict(converge_eps=2e-5)
    P, Po = both_params(**kw)
    al = Aligner(640, 480, 8, 4)
    try:
        T, st = al.align_pairs(frames[1:4], frames[0:3], intr, P)
        Tfull, stfull = al.align_pairs(frames[1:4], frames[0:3], intr, default_params())
        for i in range(3):
            To, so = O.align_pair(frames[i + 1], frames[i], intr, Po)
            assert st[i].status == 0
            assert st[i].iterations < 19 and abs(st[i].iterations - so.iterations) <= 1, (st[i].iterations, so.iterations)
            dt, dr = synth.pose_error(T[i], To)
            assert dt < 1e-4 and dr < 1e-4
            ft, fr = synth.pose_error(T[i], Tfull[i])               # stopping early costs (almost) nothing
            assert ft < 1e-4 and fr < 1e-4
            assert stfull[i].iterations == 19
    finally:
        al.close()
