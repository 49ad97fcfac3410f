"""Photometric term (BASELINE config 4): Oracle-N on CPU, CUDA parity on the GPU.
The reference's PhotometricCost is dead code (photometric_cost.hpp:48-65 calls undefined functions), so
only its conventions are inherited: residual sign I_dst(pi(T p)) - I_src (:63), bilinear sampling clamped at
the borders (sample.hpp:32-47)."""
import numpy as np
import pytest

from conftest import has_gpu
from oracle import oracle as O
from realsensetracker_b200 import synth

W, H, INTR = 160, 120, (96.0, 96.0, 80.0, 60.0)


@pytest.fixture(scope="module")
def rgbd():
    sc = synth.Scene(7)
    Twc = synth.trajectory(3, seed=7, step_t=0.02, step_r=0.015)
    fr = [sc.render(Twc[k], W, H, intr=INTR, rgb=True) for k in range(3)]
    depth = np.stack([f[0] for f in fr]); rgb = np.stack([f[1] for f in fr])
    gt = np.stack([synth.relative_pose(Twc[k], Twc[k + 1]) for k in range(2)])
    return depth, rgb, gt


def test_intensity_pyramid_against_numpy(rgbd):
    depth, rgb, gt = rgbd
    I0 = O.intensity(rgb[0])
    want = (0.299 * rgb[0][..., 0] + 0.587 * rgb[0][..., 1] + 0.114 * rgb[0][..., 2]) / 255.0
    assert np.allclose(I0, want, atol=2e-7) and I0.min() >= 0 and I0.max() <= 1
    I1 = O.intensity_down(I0)
    want1 = I0.reshape(H // 2, 2, W // 2, 2).astype(np.float64).mean(axis=(1, 3))
    assert I1.shape == (H // 2, W // 2) and np.allclose(I1, want1, atol=1e-7)


def test_photometric_normal_equations_against_numpy(rgbd):
    depth, rgb, gt = rgbd
    P = O.default_params(photo_weight=0.5)
    L = O.level_info(INTR, W, H, 0)
    G = O.geometry(depth[0], L, P)
    Is, Id = O.intensity(rgb[1]), O.intensity(rgb[0])
    # not the identity: there u_f is an exact integer and fp32/fp64 floor() pick different bilinear cells
    T = (gt[0] @ synth.make_pose(synth.rotvec_to_R([0.004, -0.003, 0.002]), [0.003, 0.002, -0.004])).astype(np.float32).astype(np.float64)
    idx, st_geo = O.evaluate(depth[1], None, G, L, P, T)
    idx2, st = O.evaluate_photo(depth[1], None, G, Is, Id, L, P, T)
    assert np.array_equal(idx, idx2) and st.count == st_geo.count
    # float64 restatement of the photometric rows only
    fx, fy, cx, cy = INTR
    v, u = np.nonzero(idx >= 0)
    z = depth[1][v, u] * 0.001
    p = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], -1)
    q = p @ T[:3, :3].T + T[:3, 3]
    uf, vf = fx * q[:, 0] / q[:, 2] + cx, fy * q[:, 1] / q[:, 2] + cy
    x0, y0 = np.floor(uf).astype(int), np.floor(vf).astype(int)
    ax, ay = uf - x0, vf - y0
    c = lambda a, hi: np.clip(a, 0, hi)
    I = Id.astype(np.float64)
    I00, I10 = I[c(y0, H - 1), c(x0, W - 1)], I[c(y0, H - 1), c(x0 + 1, W - 1)]
    I01, I11 = I[c(y0 + 1, H - 1), c(x0, W - 1)], I[c(y0 + 1, H - 1), c(x0 + 1, W - 1)]
    top, bot = I00 + ax * (I10 - I00), I01 + ax * (I11 - I01)
    val, gv, gu = top + ay * (bot - top), bot - top, (I10 - I00) + ay * ((I11 - I01) - (I10 - I00))
    r = np.sqrt(0.5) * (val - Is[v, u])
    d = np.sqrt(0.5) * np.stack([gu * fx / q[:, 2], gv * fy / q[:, 2], -(gu * fx * q[:, 0] + gv * fy * q[:, 1]) / q[:, 2] ** 2], -1)
    J = np.concatenate([np.cross(q, d), d], 1)
    dA = np.array(st.A[:]) - np.array(st_geo.A[:])
    db = np.array(st.b[:]) - np.array(st_geo.b[:])
    assert np.allclose(dA, (J.T @ J)[np.triu_indices(6)], rtol=2e-3, atol=2e-3 * np.abs(dA).max())
    assert np.allclose(db, J.T @ r, rtol=2e-2, atol=2e-3 * np.abs(db).max())
    assert np.isclose(st.sum_wr2 - st_geo.sum_wr2, (r * r).sum(), rtol=1e-3)


def test_texture_resolves_a_geometrically_degenerate_scene():
    """A camera facing one flat wall: geometry cannot see in-plane translation; the texture can."""
    sc = synth.Scene(0)
    sc._s.n_spheres = 0
    sc._s.n_boxes = 0
    intr = (400.0, 400.0, 80.0, 60.0)                       # narrow field of view: only the far wall is visible
    T_src = synth.make_pose(np.eye(3), [0.02, -0.015, 0.0])  # pure in-plane shift
    d0, c0 = sc.render(np.eye(4), W, H, intr=intr, rgb=True)
    d1, c1 = sc.render(T_src, W, H, intr=intr, rgb=True)
    assert d0.min() == d0.max() == 4000
    gt = synth.relative_pose(np.eye(4), T_src)
    Pg = O.default_params(damping=1e-3)                      # damping keeps the rank-deficient geometric system solvable
    Pp = O.default_params(damping=1e-3, photo_weight=10.0)
    Tg, sg = O.align_pair_rgbd(d1, d0, c1, c0, intr, Pg)
    Tp, sp = O.align_pair_rgbd(d1, d0, c1, c0, intr, Pp)
    eg, ep = synth.pose_error(Tg, gt), synth.pose_error(Tp, gt)
    assert eg[0] > 0.02                                      # geometry alone: the shift is not recovered
    assert ep[0] < 0.004 and ep[1] < 0.004, ep               # with the photometric term it is


@pytest.mark.gpu
def test_gpu_photometric_parity(rgbd):
    from realsensetracker_b200 import Aligner, default_params
    depth, rgb, gt = rgbd
    kw = dict(photo_weight=0.5)
    P, Po = default_params(**kw), O.default_params(**kw)
    al = Aligner(W, H, 6, 3)
    try:
        al.begin(W, H, INTR, P)
        al.upload(depth, rgb=rgb)
        al.preprocess(0, 3)
        I = O.intensity(rgb[1])
        for l in range(3):
            assert np.array_equal(al.read_intensity(1, l).view(np.uint32), I.view(np.uint32)), f"intensity level {l}"
            I = O.intensity_down(I)
        # one evaluation: association bit-exact, normal equations within 1e-4
        d_s, d_d, Is, Id = depth[1], depth[0], O.intensity(rgb[1]), O.intensity(rgb[0])
        for l in range(3):
            L = O.level_info(INTR, W, H, l)
            if l > 0:
                d_s, d_d = O.pyr_down(d_s, Po.pyr_depth_tol), O.pyr_down(d_d, Po.pyr_depth_tol)
                Is, Id = O.intensity_down(Is), O.intensity_down(Id)
            idx_o, st_o = O.evaluate_photo(d_s, None, O.geometry(d_d, L, Po), Is, Id, L, Po, np.eye(4))
            idx_g, st_g = al.evaluate(1, 0, l, np.eye(4))
            assert np.array_equal(idx_g, idx_o)
            Ao, Ag = np.array(st_o.A[:]), np.array(st_g.A[:])
            assert np.max(np.abs(Ag - Ao)) <= 1e-4 * np.max(np.abs(Ao))
            assert abs(st_g.sum_wr2 - st_o.sum_wr2) <= 1e-4 * st_o.sum_wr2
        # full alignment, frame-to-keyframe style: frames 1 and 2 both against keyframe 0
        Tg, st = al.align_pairs(depth[[1, 2]], depth[[0, 0]], INTR, P, src_rgb=rgb[[1, 2]], dst_rgb=rgb[[0, 0]])
        for i, k in enumerate((1, 2)):
            To, so = O.align_pair_rgbd(depth[k], depth[0], rgb[k], rgb[0], INTR, Po)
            dt, dr = synth.pose_error(Tg[i], To)
            assert st[i].status == 0 and dt < 1e-4 and dr < 1e-4, (dt, dr)
        gt02 = gt[0] @ gt[1]
        assert synth.pose_error(Tg[1], gt02)[0] < 5e-3
        with pytest.raises(Exception):                       # photo_weight > 0 without rgb is an error, not a silent fallback
            al.align_pairs(depth[1:2], depth[0:1], INTR, P)
    finally:
        al.close()
