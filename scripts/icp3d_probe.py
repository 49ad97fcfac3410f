"""Experiment: where does rst_icp3d_depth spend its time (CTAs per pair, grid cell size, batch)?"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from realsensetracker_b200 import Aligner, synth

W, H = 640, 480
intr = synth.intrinsics_for(W, H)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 129
frames, gt = synth.render_sequence(n, W, H, seed=0)
al = Aligner(16, 16, 2, 1)
s, d = np.arange(1, n, dtype=np.int32), np.arange(0, n - 1, dtype=np.int32)
al.icp3d_depth(frames, s, d, intr, max_iter=1)


def run(np_, cell, it, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        ok, T, mc, cnt = al.icp3d_depth(frames[:np_ + 1], s[:np_], d[:np_], intr, max_iter=it, grid_cell=cell)
        best = min(best, time.perf_counter() - t0)
    err = max(synth.pose_error(T[i], gt[i])[0] for i in range(np_))
    return best, cnt.mean(), err


for np_ in (1, 8, n - 1):
    for c in (1, 2, 4, 8, 16, 0):
        al.set_icp3d_cluster(c)
        for cell in (0.1,) if c else (0.06, 0.1, 0.2):
            dt, pts, err = run(np_, cell, 128)
            print(f"pairs {np_:3d} ctas/pair {c:2d} cell {cell:4.2f}: {dt*1e3:8.2f} ms  ({np_/dt:8.0f} pairs/s)  mean pts {pts:.0f}  err_t {err:.4f}")
