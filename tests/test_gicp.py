"""GICP plane-to-plane: ComputeCovariances (point_cloud_utils.cpp:100-161) and the GICP residual
with ceres::HuberLoss(0.5) (gicp_cost.hpp:40-73, align_gicp.cpp:59-77). CPU: the numpy restatement against the
reference's own ComputeCovariances (compiled unmodified, oracle/_ref) and against finite differences. GPU: the CUDA
residual / cost / normal equations against the restatement, and the alignment loop on depth-derived clouds."""
import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O
from realsensetracker_b200 import synth

GN = np.load(ROOT / "tests" / "golden" / "oracle_n_160x120.npz")


def clouds():
    intr = tuple(GN["intr"])
    s = O.downsample_voxel(O.remove_nans(O.backproject(GN["frames"][1], intr)), 0.05)
    d = O.downsample_voxel(O.remove_nans(O.backproject(GN["frames"][0], intr)), 0.05)
    s = s[np.abs(s).sum(1) > 0]; d = d[np.abs(d).sum(1) > 0]          # drop the invalid-depth point at the origin
    return s, d


@pytest.mark.skipif(O.ref_lib() is None, reason="needs oracle/_ref/libref.so")
def test_covariance_restatement_matches_the_compiled_reference():
    s, _ = clouds()
    for gicp in (False, True):
        ref = O.ref_covariances(s, gicp).astype(np.float64)
        mine = O.covariances(s, gicp)
        err = np.abs(ref - mine).max(axis=(1, 2)) / np.abs(ref).max(axis=(1, 2))
        if not gicp:
            assert err.max() < 1e-3
        else:
            assert np.median(err) < 1e-4 and (err < 1e-2).mean() > 0.97


def test_gicp_restatement_is_self_consistent():
    s, d = clouds()
    Cs, Cd = O.covariances(s, True), O.covariances(d, True)
    idx, _ = O.nn(d, s)
    T = synth.make_pose(synth.rotvec_to_R([0.004, -0.003, 0.002]), [0.003, 0.002, -0.004])
    r = O.gicp_evaluate(s, d, Cs, Cd, idx, T, huber=0.5)
    e = r["residuals"]
    sq = (e * e).sum(1)
    rho = np.where(sq > 0.25, 2 * 0.5 * np.sqrt(sq) - 0.25, sq)          # ceres::HuberLoss(0.5): rho(s) = 2 a sqrt(s) - a^2 beyond a^2
    assert np.isclose(r["cost"], 0.5 * rho.sum()) and r["count"] == len(s)
    # whitening: e^T e = delta^T C^-1 delta
    R, t = T[:3, :3], T[:3, 3]
    i = 17
    delta = R @ s[i].astype(np.float64) + t - d[idx[i]].astype(np.float64)
    C = Cd[idx[i]] + R @ Cs[i] @ R.T
    assert np.isclose(sq[i], delta @ np.linalg.solve(C, delta), rtol=1e-9)
    # Jacobian of the whitened residual w.r.t. a left perturbation (C held at the current rotation): finite differences
    J = r["J"][i]
    from scipy.linalg import sqrtm, expm
    Wh = np.real(np.linalg.inv(sqrtm(C)))
    for k in range(6):
        xi = np.zeros(6); xi[k] = 1e-6
        Om = np.array([[0, -xi[2], xi[1]], [xi[2], 0, -xi[0]], [-xi[1], xi[0], 0]])
        p1 = expm(Om) @ (R @ s[i] + t) + xi[3:]
        num = (Wh @ (p1 - d[idx[i]]) - Wh @ (R @ s[i] + t - d[idx[i]])) / 1e-6
        assert np.allclose(J[:, k], num, atol=1e-4 * max(1.0, np.abs(num).max()))


@pytest.mark.gpu
def test_gpu_gicp_residuals_cost_and_normal_equations():
    from realsensetracker_b200 import Aligner
    s, d = clouds()
    al = Aligner(16, 16, 2, 1)
    try:
        T = synth.make_pose(synth.rotvec_to_R([0.004, -0.003, 0.002]), [0.003, 0.002, -0.004]).astype(np.float32).astype(np.float64)
        idx, _ = al.find_correspondences(d, (s @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
        for gicp_cov, tol in ((True, 1e-4), (False, 5e-3)):
            Cs, Cd = al.cloud_covariances(s, gicp_cov), al.cloud_covariances(d, gicp_cov)
            want = O.gicp_evaluate(s, d, Cs, Cd, idx, T, huber=0.5)
            res, st = al.gicp_evaluate(s, d, Cs, Cd, idx, T, huber=0.5)
            assert st.count == want["count"] == len(s)
            scale = np.abs(want["residuals"]).max()
            err = np.abs(res - want["residuals"]).max(axis=1) / scale
            assert np.quantile(err, 0.99) < tol, (gicp_cov, np.quantile(err, 0.99))
            assert abs(st.cost - want["cost"]) <= 10 * tol * want["cost"]
            A, Ao = np.array(st.A[:]), want["A"]
            assert np.max(np.abs(A - Ao)) <= 10 * tol * np.max(np.abs(Ao))
            b, bo = np.array(st.b[:]), want["b"]
            assert np.max(np.abs(b - bo) / np.sqrt(Ao[[0, 6, 11, 15, 18, 20]] * 2 * want["cost"])) <= 10 * tol
        # no robust loss: cost = 1/2 sum |e|^2
        res, st = al.gicp_evaluate(s, d, Cs, Cd, idx, T, huber=0.0)
        assert np.isclose(st.cost, 0.5 * (res.astype(np.float64) ** 2).sum(), rtol=1e-5)
    finally:
        al.close()


@pytest.mark.gpu
def test_gpu_gicp_alignment_recovers_the_motion():
    from realsensetracker_b200 import Aligner
    s, d = clouds()
    gt = GN["gt"][0]
    al = Aligner(16, 16, 2, 1)
    try:
        T, st = al.gicp_align(s, d, max_outer=16, inner_iters=4, huber=0.5, use_gicp_covariances=True)
        et, er = synth.pose_error(T, gt)
        ok_i, T_i = al.icp3d_pairs([s], [d], 128)
        it, ir = synth.pose_error(T_i[0], gt)
        print(f"gicp {et:.4f} m {er:.4f} rad (cost {st.cost:.3f}, {st.count} pairs); point-to-point icp {it:.4f} m {ir:.4f} rad")
        assert st.count == len(s) and np.isfinite(st.cost)
        assert et < 0.02 and er < 0.02
        assert et < it + 2e-3                                   # plane-to-plane is not worse than the reference's point-to-point ICP
        T2, _ = al.gicp_align(s, d, max_outer=16, inner_iters=4, huber=0.5, use_gicp_covariances=True)
        assert np.array_equal(T, T2)                            # deterministic
    finally:
        al.close()


@pytest.mark.gpu
def test_gpu_gicp_minimize_over_fixed_correspondences():
    """The 7-argument ComputeAlignment (align_gicp.cpp:41-117): for given covariances and correspondences the returned
    pose lowers the cost of the seed, is a stationary point of it (the Gauss-Newton step there is negligible), its
    statistics are those of rst_gicp_evaluate at that pose, and 16 rounds of { FindCorrespondences; minimise } driven
    from the host reproduce the 3-argument form."""
    from realsensetracker_b200 import Aligner
    s, d = clouds()
    al = Aligner(16, 16, 2, 1)

    def gn_step(st):
        A = np.zeros((6, 6)); A[np.triu_indices(6)] = np.array(st.A[:]); A = A + A.T - np.diag(np.diag(A))
        return np.linalg.solve(A, -np.array(st.b[:]))

    try:
        Cs, Cd = al.cloud_covariances(s, True), al.cloud_covariances(d, True)
        idx, _ = al.find_correspondences(d, s)
        I = np.eye(4)
        _, st0 = al.gicp_evaluate(s, d, Cs, Cd, idx, I, huber=0.5, want_residuals=False)
        T0, stz = al.gicp_minimize(s, d, Cs, Cd, idx, T0=I, max_iters=0)
        assert np.array_equal(T0, I) and stz.cost == st0.cost and stz.count == st0.count == len(s)
        T1, st1 = al.gicp_minimize(s, d, Cs, Cd, idx, T0=I, max_iters=32)
        assert st1.count == len(s) and st1.cost < st0.cost
        _, st_at = al.gicp_evaluate(s, d, Cs, Cd, idx, T1, huber=0.5, want_residuals=False)
        assert np.isclose(st_at.cost, st1.cost, rtol=1e-6)
        x0, x1 = np.linalg.norm(gn_step(st0)), np.linalg.norm(gn_step(st_at))
        print(f"gicp minimise: cost {st0.cost:.4f} -> {st1.cost:.4f}, Gauss-Newton step {x0:.2e} -> {x1:.2e}")
        assert x1 < 1e-3 and x1 < 0.2 * x0
        # against the float64 CPU restatement of the same minimisation (same covariances and correspondences)
        To, cost_o, _ = O.gicp_minimize(s, d, Cs, Cd, idx, T0=I, max_iters=32, huber=0.5)
        assert abs(st1.cost - cost_o) <= 1e-3 * cost_o, (st1.cost, cost_o)
        dt, dr = synth.pose_error(T1, To)
        assert dt < 1e-3 and dr < 1e-3, (dt, dr)
        T2, st2 = al.gicp_minimize(s, d, Cs, Cd, idx, T0=I, max_iters=32)
        assert np.array_equal(T1, T2) and st2.cost == st1.cost             # deterministic
        # the 3-argument form = this call inside the correspondence loop
        T = I.copy()
        for _ in range(16):
            idx_k, _ = al.find_correspondences(d, (s.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
            T, _ = al.gicp_minimize(s, d, Cs, Cd, idx_k, T0=T, max_iters=4)
        Ta, _ = al.gicp_align(s, d, max_outer=16, inner_iters=4, huber=0.5, use_gicp_covariances=True)
        dt, dr = synth.pose_error(T, Ta)
        assert dt < 1e-3 and dr < 1e-3, (dt, dr)
    finally:
        al.close()


def test_gicp_minimize_restatement_and_the_recorded_gpu_run():
    """Oracle: Levenberg-Marquardt over gicp_evaluate's normal equations reaches a stationary point of the Huber GICP cost
    on the golden clouds, and the costs are the ones the CUDA path printed on the B200
    (profiles/r02_gicp_minimize_gpu_test.log: the -s output of test_gpu_gicp_minimize_over_fixed_correspondences)."""
    import re
    s, d = clouds()
    Cs, Cd = O.covariances(s, True), O.covariances(d, True)
    idx, _ = O.nn(d, s)
    T, cost, hist = O.gicp_minimize(s, d, Cs, Cd, idx, max_iters=32)
    assert all(b <= a for a, b in zip(hist, hist[1:])) and cost < 0.1 * hist[0]
    r = O.gicp_evaluate(s, d, Cs, Cd, idx, T, huber=0.5)
    A = np.zeros((6, 6)); A[np.triu_indices(6)] = r["A"]; A = A + A.T - np.diag(np.diag(A))
    assert np.linalg.norm(np.linalg.solve(A, -r["b"])) < 1e-7
    et, er = synth.pose_error(T, GN["gt"][0])
    assert et < 5e-3 and er < 5e-3                               # one round of correspondences from the identity
    log = (ROOT / "profiles" / "r02_gicp_minimize_gpu_test.log").read_text()
    m = re.search(r"gicp minimise: cost ([0-9.]+) -> ([0-9.]+)", log)
    assert m, "the recorded GPU line is missing"
    assert abs(float(m.group(1)) - hist[0]) <= 2e-4 * hist[0] and abs(float(m.group(2)) - cost) <= 2e-4 * cost
