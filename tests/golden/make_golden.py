#!/usr/bin/env python
"""Generates tests/golden/*.npz — regression vectors for the oracles and the CUDA path.

The reference (yycho0108/RealsenseTracker) holds no golden vectors, tests or fixtures and cannot be
built or imported here, so these vectors pin OUR specification (Oracle-N) and OUR restatement of the
reference algorithm (Oracle-R) against accidental change; they are not reference outputs
("parity unpinned"). Inputs are regenerated from the seeded synthetic source and
stored too, so the fixture does not depend on the renderer staying bit-stable.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402
from realsensetracker_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    w, h = 160, 120
    intr = (96.0, 96.0, 80.0, 60.0)
    scene = synth.Scene(7)
    Twc = synth.trajectory(3, seed=7, step_t=0.02, step_r=0.015)
    frames = np.stack([scene.render(Twc[k], w, h, intr=intr) for k in range(3)])
    gt = np.stack([synth.relative_pose(Twc[k], Twc[k + 1]) for k in range(2)])

    P = O.default_params()
    L0 = O.level_info(intr, w, h, 0)
    d1 = O.pyr_down(frames[0], P.pyr_depth_tol)
    d2 = O.pyr_down(d1, P.pyr_depth_tol)
    G0 = O.geometry(frames[0], L0, P)
    idx, st = O.evaluate(frames[1], None, G0, L0, P, np.eye(4))
    T, sa = O.align_pair(frames[1], frames[0], intr, P)
    Ph = O.default_params(robust_kind=1, robust_scale=0.002, normal_cos_min=0.9)
    Gs0 = O.geometry(frames[1], L0, Ph)
    idx_h, st_h = O.evaluate(frames[1], Gs0, O.geometry(frames[0], L0, Ph), L0, Ph, np.eye(4))
    np.savez_compressed(
        OUT / "oracle_n_160x120.npz", frames=frames, gt=gt, intr=np.array(intr),
        pyr1=d1, pyr2=d2, G0=G0, idx=idx, A=np.array(st.A[:]), b=np.array(st.b[:]), count=st.count,
        sum_wr2=st.sum_wr2, pose=T, pose_rmse=sa.rmse, pose_count=sa.count, pose_A=np.array(sa.A[:]),
        idx_huber_ngate=idx_h, A_huber_ngate=np.array(st_h.A[:]), count_huber_ngate=st_h.count)

    # Oracle-R: the reference's disabled self-test, made live (rs_align_app.cpp:257-263):
    # dst = R_x(0.1) R_y(-0.2) R_z(0.25) * src (+ a translation), recover it.
    rng = np.random.default_rng(11)
    src = rng.uniform(-1, 1, size=(600, 3)).astype(np.float32)        # RandomSource-style cloud (data_source.hpp:29-36)
    Tk = synth.make_pose(synth.rot_xyz(0.1, -0.2, 0.25) , [0.05, -0.03, 0.02])
    Tk_small = synth.make_pose(synth.rot_xyz(0.02, -0.04, 0.05), [0.05, -0.03, 0.02])
    dst = (src @ Tk_small[:3, :3].T + Tk_small[:3, 3]).astype(np.float32)
    ok, Tr, ex = O.align_icp3d(src, dst, 128, details=True)
    pairs = np.stack([np.arange(600), np.arange(600)], 1).astype(np.int32)
    dst_big = (src @ Tk[:3, :3].T + Tk[:3, 3]).astype(np.float32)
    okk, Tkab = O.solve_kabsch(src, dst_big, pairs)
    wts = rng.uniform(0.1, 1.0, 600).astype(np.float32)
    okw, Tkab_w = O.solve_kabsch(src, dst_big, pairs, wts)
    vox = O.downsample_voxel(src, 0.25)
    okd, Td, exd = O.align_depth_pair(frames[1], frames[0], intr)
    np.savez_compressed(
        OUT / "oracle_r.npz", src=src, dst=dst, dst_big=dst_big, T_small=Tk_small, T_big=Tk, icp_T=Tr,
        icp_mean_cost=ex["mean_cost"], icp_nbrs=ex["nbrs"], icp_weights=ex["weights"], icp_cov=ex["cov"],
        kabsch_T=Tkab, kabsch_w=wts, kabsch_T_weighted=Tkab_w, voxel_025=vox,
        depth_pair_T=Td, depth_pair_mean_cost=exd["mean_cost"], depth_pair_n=np.array([exd["n_src"], exd["n_dst"]]))
    print("wrote", [p.name for p in OUT.glob("*.npz")])


if __name__ == "__main__":
    main()
