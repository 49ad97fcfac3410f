"""CPU: the synthetic frame source is deterministic, thread-count independent and geometrically consistent."""
import os

import numpy as np

from realsensetracker_b200 import synth


def test_render_is_deterministic_and_in_range():
    s = synth.Scene(0)
    a = s.render(np.eye(4), 160, 120, intr=(96.0, 96.0, 80.0, 60.0))
    b = synth.Scene(0).render(np.eye(4), 160, 120, intr=(96.0, 96.0, 80.0, 60.0))
    assert np.array_equal(a, b) and a.dtype == np.uint16 and a.shape == (120, 160)
    assert (a > 0).all() and a.max() <= 4100 and a.min() >= 500      # room depth range, camera inside
    assert a[60, 80] == 4000                                         # optical axis hits the far wall at z = 4 m


def test_noise_and_invalid_injection_are_seeded():
    s = synth.Scene(2)
    nz = synth.Noise(sigma_lsb_at_1m=1.0, p_invalid_pixel=0.2, p_invalid_block=0.1)
    a = s.render(np.eye(4), 160, 120, noise=nz, frame_seed=5)
    b = s.render(np.eye(4), 160, 120, noise=nz, frame_seed=5)
    c = s.render(np.eye(4), 160, 120, noise=nz, frame_seed=6)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert 0.2 < (a == 0).mean() < 0.4


def test_reprojection_consistency_of_ground_truth():
    """A point back-projected from frame k+1 and moved by the ground-truth pose lands on frame k's surface."""
    frames, gt = synth.render_sequence(2, 160, 120, seed=4)
    fx, fy, cx, cy = synth.intrinsics_for(160, 120)
    v, u = np.mgrid[10:110:7, 10:150:7]
    z = frames[1][v, u] * 0.001
    p = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], -1).reshape(-1, 3)
    q = p @ gt[0][:3, :3].T + gt[0][:3, 3]
    ui = np.rint(fx * q[:, 0] / q[:, 2] + cx).astype(int)
    vi = np.rint(fy * q[:, 1] / q[:, 2] + cy).astype(int)
    ok = (ui >= 0) & (ui < 160) & (vi >= 0) & (vi < 120)
    dz = np.abs(frames[0][vi[ok], ui[ok]] * 0.001 - q[ok, 2])
    assert np.median(dz) < 0.01


def test_rgb_output_and_pairs_api():
    s = synth.Scene(1)
    d, rgb = s.render(np.eye(4), 64, 48, rgb=True)
    assert rgb.shape == (48, 64, 3) and rgb.dtype == np.uint8 and rgb.std() > 5
    src, dst, gt = synth.render_pairs(2, 64, 48, seed=3)
    assert src.shape == dst.shape == (2, 48, 64) and gt.shape == (2, 4, 4)
    assert np.allclose(gt[:, :3, :3] @ gt[:, :3, :3].transpose(0, 2, 1), np.eye(3), atol=1e-12)
