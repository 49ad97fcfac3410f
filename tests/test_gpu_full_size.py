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
