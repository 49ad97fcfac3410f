"""GPU: BASELINE.json's full sizes (configs 3, 4, 5) through size-independent properties —
known-motion round trip, inverse consistency, idempotence at the fixed point, determinism —
plus one oracle parity pair per size (the oracle needs ~1 s per 1280x720 pair)."""
import numpy as np
import pytest

from oracle import oracle as O
from realsensetracker_b200 import Aligner, default_params, synth
from realsensetracker_b200 import _native as N

pytestmark = pytest.mark.gpu


def test_1280x720_batch_properties_and_parity():
    w, h = 1280, 720
    intr = synth.intrinsics_for(w, h)
    n = 6
    src, dst, gt = synth.render_pairs(n, w, h, seed=5)                 # ||t|| <= 3 cm, angle <= 2 deg (config 3)
    P, Po = default_params(), O.default_params()
    al = Aligner(w, h, 2 * n, n)
    try:
        T, st = al.align_pairs(src, dst, intr, P)
        T2, _ = al.align_pairs(src, dst, intr, P)
        assert np.array_equal(T, T2)                                    # deterministic
        Tinv, _ = al.align_pairs(dst, src, intr, P)                     # the opposite direction
        Tfix, _ = al.align_pairs(src, dst, intr, default_params(num_levels=1, iters=[3, 0, 0, 0]), T0=T)
        for i in range(n):
            assert st[i].status == 0 and st[i].count > 0.8 * w * h
            et, er = synth.pose_error(T[i], gt[i])
            assert et < 1.5e-3 and er < 1.5e-3, (i, et, er)             # known motion recovered
            ct, cr = synth.pose_error(T[i] @ Tinv[i], np.eye(4))
            assert ct < 1e-3 and cr < 1e-3, (i, ct, cr)                 # inverse consistency
            ft, fr = synth.pose_error(Tfix[i], T[i])
            assert ft < 5e-5 and fr < 5e-5, (i, ft, fr)                 # converged pose is a fixed point
        T1, _ = al.align_pairs(src[3:4], dst[3:4], intr, P)
        assert np.array_equal(T1[0], T[3])                              # independent of the batch
        To, so = O.align_pair(src[0], dst[0], intr, Po)
        dt, dr = synth.pose_error(T[0], To)
        assert dt < 1e-4 and dr < 1e-4 and abs(st[0].count - so.count) <= 8
    finally:
        al.close()


def test_848x480_levels_with_row_tails():
    """848 -> 424 -> 212: the coarsest width is not a multiple of 8 (pitch padding) nor of 64 (chunk tails)."""
    w, h = 848, 480
    intr = synth.intrinsics_for(w, h)
    frames, gt = synth.render_sequence(3, w, h, seed=2)
    P, Po = default_params(), O.default_params()
    al = Aligner(w, h, 3, 2)
    try:
        al.begin(w, h, intr, P)
        al.upload(frames)
        al.preprocess(0, 3)
        d = frames[0]
        for l in range(3):
            L = O.level_info(intr, w, h, l)
            if l > 0:
                d = O.pyr_down(d, Po.pyr_depth_tol)
            assert al.level_info(l)[:2] == (L.w, L.h)
            assert np.array_equal(al.read_depth(0, l), d)
            assert np.array_equal(al.read_geometry(0, l).view(np.uint32), O.geometry(d, L, Po).view(np.uint32))
        L2 = O.level_info(intr, w, h, 2)
        d1 = O.pyr_down(O.pyr_down(frames[1], 100), 100)
        idx_o, st_o = O.evaluate(d1, None, O.geometry(d, L2, Po), L2, Po, np.eye(4))
        idx_g, st_g = al.evaluate(1, 0, 2, np.eye(4))
        assert np.array_equal(idx_g, idx_o) and st_g.count == st_o.count
        T, st = al.align_sequence(frames, intr, P)
        for i in range(2):
            To, _ = O.align_pair(frames[i + 1], frames[i], intr, Po)
            dt, dr = synth.pose_error(T[i], To)
            assert dt < 1e-4 and dr < 1e-4
            assert synth.pose_error(T[i], gt[i])[0] < 2e-3
    finally:
        al.close()


def test_stress_invalid_depth_large_motion_huber():
    """Config 5: 30 % invalid depth (pixels + 16x16 blocks), motion up to 10 cm / 8 deg, Huber weights."""
    w, h = 640, 480
    intr = synth.intrinsics_for(w, h)
    noise = synth.Noise(p_invalid_pixel=0.21, p_invalid_block=0.12)
    src, dst, gt = synth.render_pairs(4, w, h, seed=9, max_t=0.10, max_r=np.deg2rad(8.0), noise=noise)
    assert 0.25 < (src == 0).mean() < 0.36
    kw = dict(robust_kind=N.RST_ROBUST_HUBER, robust_scale=0.01, dist_max=0.6, iters=[10, 10, 12, 0], num_levels=3)
    P, Po = default_params(**kw), O.default_params(**kw)
    al = Aligner(w, h, 8, 4)
    try:
        T, st = al.align_pairs(src, dst, intr, P)
        for i in range(4):
            et, er = synth.pose_error(T[i], gt[i])
            assert st[i].status == 0 and et < 5e-3 and er < 5e-3, (i, et, er, synth.pose_error(np.eye(4), gt[i]))
        To, so = O.align_pair(src[1], dst[1], intr, Po)
        dt, dr = synth.pose_error(T[1], To)
        assert dt < 1e-4 and dr < 1e-4, (dt, dr)
    finally:
        al.close()


def test_config4_848x480_rgbd_frame_to_keyframe_photometric_parity():
    """BASELINE config 4 at its own size: 848x480 RGB-D, frames 1..3 against keyframe 0, geometric + photometric
    residual (lambda = 0.5). Intensity pyramid and association bit-exact, normal equations 1e-4 at the finest level,
    final poses within 1e-4 m / 1e-4 rad of the CPU specification, known motion recovered."""
    w, h = 848, 480
    intr = synth.intrinsics_for(w, h)
    sc = synth.Scene(4)
    Twc = synth.trajectory(4, seed=4)
    fr = [sc.render(Twc[k], w, h, intr=intr, rgb=True) for k in range(4)]
    depth = np.stack([f[0] for f in fr]); rgb = np.stack([f[1] for f in fr])
    kw = dict(photo_weight=0.5)
    P, Po = default_params(**kw), O.default_params(**kw)
    al = Aligner(w, h, 6, 3)
    try:
        al.begin(w, h, intr, P)
        al.upload(depth, rgb=rgb)
        al.preprocess(0, 4)
        I = O.intensity(rgb[2])
        for l in range(3):
            assert np.array_equal(al.read_intensity(2, l).view(np.uint32), I.view(np.uint32)), f"intensity level {l}"
            I = O.intensity_down(I)
        L0 = O.level_info(intr, w, h, 0)
        T = np.eye(4)
        idx_o, st_o = O.evaluate_photo(depth[2], None, O.geometry(depth[0], L0, Po), O.intensity(rgb[2]), O.intensity(rgb[0]), L0, Po, T)
        idx_g, st_g = al.evaluate(2, 0, 0, T)
        assert np.array_equal(idx_g, idx_o)
        Ao, Ag = np.array(st_o.A[:]), np.array(st_g.A[:])
        assert np.max(np.abs(Ag - Ao)) <= 1e-4 * np.max(np.abs(Ao))
        assert abs(st_g.sum_wr2 - st_o.sum_wr2) <= 1e-4 * st_o.sum_wr2
        Tg, st = al.align_pairs(depth[[1, 2, 3]], depth[[0, 0, 0]], intr, P, src_rgb=rgb[[1, 2, 3]], dst_rgb=rgb[[0, 0, 0]])
        for i, k in enumerate((1, 2, 3)):
            gt = synth.relative_pose(Twc[0], Twc[k])
            et, er = synth.pose_error(Tg[i], gt)
            assert st[i].status == 0 and et < 2e-3 and er < 2e-3, (k, et, er)
        To, so = O.align_pair_rgbd(depth[3], depth[0], rgb[3], rgb[0], intr, Po)
        dt, dr = synth.pose_error(Tg[2], To)
        assert dt < 1e-4 and dr < 1e-4, (dt, dr)
    finally:
        al.close()


def test_config5_1280x720_huber_normal_gate_invalid_depth_parity():
    """BASELINE config 5 ingredients at 1280x720: 30 % invalid depth, Huber weights and the source/target normal
    gate. Association bit-exact and normal equations 1e-4 on all three levels, final pose 1e-4 vs the specification."""
    w, h = 1280, 720
    intr = synth.intrinsics_for(w, h)
    noise = synth.Noise(p_invalid_pixel=0.21, p_invalid_block=0.12)
    src, dst, gt = synth.render_pairs(2, w, h, seed=13, max_t=0.06, max_r=np.deg2rad(4.0), noise=noise)
    kw = dict(robust_kind=N.RST_ROBUST_HUBER, robust_scale=0.01, normal_cos_min=0.8, dist_max=0.4, iters=[10, 8, 8, 0], num_levels=3)
    P, Po = default_params(**kw), O.default_params(**kw)
    al = Aligner(w, h, 4, 2)
    try:
        T, st = al.align_pairs(src, dst, intr, P)          # slots: dst 0..1, src 2..3
        ds, dd = src[0], dst[0]
        T0 = np.eye(4)
        for l in range(3):
            L = O.level_info(intr, w, h, l)
            if l > 0:
                ds, dd = O.pyr_down(ds, Po.pyr_depth_tol), O.pyr_down(dd, Po.pyr_depth_tol)
            idx_o, st_o = O.evaluate(ds, O.geometry(ds, L, Po), O.geometry(dd, L, Po), L, Po, T0)
            idx_g, st_g = al.evaluate(2, 0, l, T0)
            assert np.array_equal(idx_g, idx_o), (l, int((idx_g != idx_o).sum()))
            Ao, Ag = np.array(st_o.A[:]), np.array(st_g.A[:])
            assert np.max(np.abs(Ag - Ao)) <= 1e-4 * np.max(np.abs(Ao)), l
        for i in range(2):
            et, er = synth.pose_error(T[i], gt[i])
            assert st[i].status == 0 and et < 5e-3 and er < 5e-3, (i, et, er)
        To, so = O.align_pair(src[0], dst[0], intr, Po)
        dt, dr = synth.pose_error(T[0], To)
        assert dt < 1e-4 and dr < 1e-4, (dt, dr)
    finally:
        al.close()
