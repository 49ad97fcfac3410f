"""CPU: Oracle-N (the specification of the CUDA path) against independent numpy/scipy restatements,
analytic cases, the committed golden vectors, and known-motion synthetic scenes."""
import numpy as np
import pytest
from scipy.linalg import expm

from conftest import ROOT
from oracle import oracle as O
from realsensetracker_b200 import synth

GOLD = np.load(ROOT / "tests" / "golden" / "oracle_n_160x120.npz")


def np_pyr_down(d, tol):
    h, w = d.shape
    b = d[: h // 2 * 2, : w // 2 * 2].astype(np.int64).reshape(h // 2, 2, w // 2, 2).transpose(0, 2, 1, 3).reshape(h // 2, w // 2, 4)
    big = np.where(b > 0, b, 1 << 40)
    m = big.min(-1, keepdims=True)
    use = (b > 0) & (b - m <= tol)
    n = use.sum(-1)
    s = (b * use).sum(-1)
    return np.where(n > 0, (s + n // 2) // np.maximum(n, 1), 0).astype(np.uint16)


def test_pyramid_matches_numpy_and_golden():
    rng = np.random.default_rng(0)
    d = rng.integers(0, 5000, size=(37, 53)).astype(np.uint16)
    d[rng.random(d.shape) < 0.3] = 0
    for tol in (0, 50, 100, 10000):
        assert np.array_equal(O.pyr_down(d, tol), np_pyr_down(d, tol))
    assert np.array_equal(O.pyr_down(GOLD["frames"][0], 100), GOLD["pyr1"])
    assert np.array_equal(O.pyr_down(GOLD["pyr1"], 100), GOLD["pyr2"])
    assert O.pyr_down(np.zeros((2, 2), np.uint16), 100).tolist() == [[0]]


def test_level_intrinsics_pixel_centre_convention():
    L1 = O.level_info((385.0, 385.0, 320.0, 240.0), 640, 480, 1)
    assert (L1.w, L1.h) == (320, 240)
    assert L1.fx == 192.5 and L1.cx == (320.0 + 0.5) / 2 - 0.5 and L1.cy == (240.0 + 0.5) / 2 - 0.5
    assert L1.ifx == np.float32(1.0) / np.float32(192.5)


def test_geometry_normals_on_an_analytic_plane():
    """Depth of the plane n.X = d seen by a pin-hole camera: normals must equal n, oriented toward the camera."""
    w, h, intr = 96, 64, (80.0, 80.0, 47.5, 31.5)
    n = np.array([0.2, -0.1, -1.0]); n /= np.linalg.norm(n)
    dpl = -2.0                                          # plane through z ~ 2 m
    u, v = np.meshgrid(np.arange(w), np.arange(h))
    ray = np.stack([(u - intr[2]) / intr[0], (v - intr[3]) / intr[1], np.ones_like(u, dtype=float)], -1)
    z = dpl / (ray @ n)
    depth = np.round(z / 0.0001).astype(np.uint16)       # 0.1 mm LSB keeps quantisation small
    P = O.default_params(depth_scale=0.0001)
    L = O.level_info(intr, w, h, 0)
    G = O.geometry(depth, L, P)
    inner = G[1:-1, 1:-1]
    assert (inner[..., 3] > 0).all()
    assert (G[0, :, 3] == 0).all() and (G[:, 0, 3] == 0).all() and (G[-1, :, 3] == 0).all()
    assert np.allclose(np.linalg.norm(inner[..., :3], axis=-1), 1.0, atol=1e-5)
    ang = np.arccos(np.clip(inner[..., :3] @ n, -1, 1))
    assert np.median(ang) < 2e-2 and ang.max() < 0.15     # quantisation-limited
    V = ray[1:-1, 1:-1] * inner[..., 3:4]
    assert ((inner[..., :3] * V).sum(-1) <= 0).all()      # orientation rule, point_cloud_utils.cpp:210-214
    assert np.allclose(inner[..., 3], depth[1:-1, 1:-1] * np.float32(0.0001))


def test_geometry_and_association_match_golden():
    P = O.default_params()
    intr = tuple(GOLD["intr"])
    L = O.level_info(intr, 160, 120, 0)
    G0 = O.geometry(GOLD["frames"][0], L, P)
    assert np.array_equal(G0.view(np.uint32), GOLD["G0"].view(np.uint32))
    idx, st = O.evaluate(GOLD["frames"][1], None, G0, L, P, np.eye(4))
    assert np.array_equal(idx, GOLD["idx"])
    assert st.count == int(GOLD["count"]) == int((idx >= 0).sum())
    assert np.array_equal(np.array(st.A[:]), GOLD["A"]) and np.array_equal(np.array(st.b[:]), GOLD["b"])
    Ph = O.default_params(robust_kind=1, robust_scale=0.002, normal_cos_min=0.9)
    Gs = O.geometry(GOLD["frames"][1], L, Ph)
    idx_h, st_h = O.evaluate(GOLD["frames"][1], Gs, G0, L, Ph, np.eye(4))
    assert np.array_equal(idx_h, GOLD["idx_huber_ngate"])
    assert st_h.count == int(GOLD["count_huber_ngate"]) < st.count
    assert np.allclose(np.array(st_h.A[:]), GOLD["A_huber_ngate"], rtol=1e-12)


def test_evaluate_against_a_numpy_restatement():
    """Independent float64 numpy restatement of association + normal equations (tolerance, not bits)."""
    P = O.default_params()
    intr = tuple(GOLD["intr"])
    fx, fy, cx, cy = intr
    L = O.level_info(intr, 160, 120, 0)
    G = GOLD["G0"].astype(np.float64)
    src = GOLD["frames"][1]
    T = GOLD["gt"][0].astype(np.float32).astype(np.float64)
    idx, st = O.evaluate(src, None, GOLD["G0"], L, P, T)
    v, u = np.nonzero(idx >= 0)
    z = src[v, u] * 0.001
    p = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], -1)
    q = p @ T[:3, :3].T + T[:3, 3]
    ui = np.rint(fx * q[:, 0] / q[:, 2] + cx).astype(int)
    vi = np.rint(fy * q[:, 1] / q[:, 2] + cy).astype(int)
    assert (np.abs(vi * 160 + ui - idx[v, u]) == 0).mean() > 0.999   # float32 vs float64 rounding ties
    g = G[vi, ui]
    qd = np.stack([(ui - cx) / fx * g[:, 3], (vi - cy) / fy * g[:, 3], g[:, 3]], -1)
    r = (g[:, :3] * (q - qd)).sum(-1)
    J = np.concatenate([np.cross(q, g[:, :3]), g[:, :3]], 1)
    A = J.T @ J
    b = J.T @ r
    Ao = np.zeros((6, 6)); Ao[np.triu_indices(6)] = st.A[:]
    assert np.allclose(np.triu(A), Ao, rtol=2e-3, atol=1e-3 * np.abs(A).max())
    assert np.allclose(b, st.b[:], rtol=2e-2, atol=2e-3 * np.abs(b).max())


def test_solve_matches_numpy_and_reports_failures():
    rng = np.random.default_rng(3)
    M = rng.normal(size=(40, 6)); A = M.T @ M; b = rng.normal(size=6)
    P = O.default_params()
    rc, xi = O.solve(A[np.triu_indices(6)], b, 1000, P)
    assert rc == 0 and np.allclose(xi, np.linalg.solve(A, -b), rtol=1e-10)
    assert O.solve(A[np.triu_indices(6)], b, 3, P)[0] == 1                 # too few -> RST_STATUS_TOO_FEW
    A1 = np.outer(M[0], M[0])                                              # rank 1 -> degenerate
    assert O.solve(A1[np.triu_indices(6)], b, 1000, P)[0] == 2
    bad = A.copy(); bad[0, 0] = np.nan
    assert O.solve(bad[np.triu_indices(6)], b, 1000, P)[0] == 4
    Pd = O.default_params(damping=10.0)
    rc, xid = O.solve(A[np.triu_indices(6)], b, 1000, Pd)
    assert np.allclose(xid, np.linalg.solve(A + 10 * np.eye(6), -b), rtol=1e-10)


@pytest.mark.parametrize("scale", [1e-7, 1e-3, 0.3, 2.0])
def test_pose_update_is_the_se3_exponential(scale):
    rng = np.random.default_rng(5)
    xi = rng.normal(size=6) * scale
    T = synth.make_pose(synth.rotvec_to_R(rng.normal(size=3)), rng.normal(size=3))
    X = np.zeros((4, 4))
    X[:3, :3] = [[0, -xi[2], xi[1]], [xi[2], 0, -xi[0]], [-xi[1], xi[0], 0]]
    X[:3, 3] = xi[3:]
    want = expm(X) @ T
    got = O.pose_update(xi, T)
    assert np.allclose(got, want, atol=1e-12)


def test_align_recovers_known_motion_and_matches_golden():
    P = O.default_params()
    intr = tuple(GOLD["intr"])
    for i in range(2):
        T, st = O.align_pair(GOLD["frames"][i + 1], GOLD["frames"][i], intr, P)
        et, er = synth.pose_error(T, GOLD["gt"][i])
        assert st.status == 0 and st.iterations == 19
        assert et < 3e-3 and er < 3e-3, (et, er)
        if i == 0:
            assert np.array_equal(T, GOLD["pose"])
            assert st.count == int(GOLD["pose_count"]) and st.rmse == np.float32(GOLD["pose_rmse"])


def test_initial_pose_convention_and_failure():
    """Pose in = initial guess, out = result, src -> dst (align_icp.cpp:82,107,156); too few points -> failure."""
    P = O.default_params(iters=[2, 0, 0], num_levels=1)
    intr = tuple(GOLD["intr"])
    f = GOLD["frames"]
    T_from_gt, _ = O.align_pair(f[1], f[0], intr, P, T0=GOLD["gt"][0])
    T_from_id, _ = O.align_pair(f[1], f[0], intr, P)
    assert synth.pose_error(T_from_gt, GOLD["gt"][0])[0] < synth.pose_error(T_from_id, GOLD["gt"][0])[0]
    zero = np.zeros_like(f[0])
    T, st = O.align_pair(zero, f[0], intr, P)
    assert st.status == 1 and np.array_equal(T, np.eye(4)) and st.count == 0


def test_convergence_early_exit_in_the_oracle():
    intr = tuple(GOLD["intr"])
    f = GOLD["frames"]
    T_full, s_full = O.align_pair(f[1], f[0], intr, O.default_params())
    T_early, s_early = O.align_pair(f[1], f[0], intr, O.default_params(converge_eps=5e-5))
    assert s_full.iterations == 19 and 3 <= s_early.iterations < 19
    dt, dr = synth.pose_error(T_early, T_full)
    assert dt < 2e-4 and dr < 2e-4
