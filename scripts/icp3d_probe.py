"""Experiment: where does rst_icp3d_depth spend its time (grid cell size, iteration count, batch)?"""
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
for cell in (0.05, 0.1, 0.2, 0.4):
    for it in (0, 1, 16, 128):
        t0 = time.perf_counter()
        ok, T, mc, cnt = al.icp3d_depth(frames, s, d, intr, max_iter=it, grid_cell=cell)
        dt = time.perf_counter() - t0
        print(f"cell {cell:4.2f} iters {it:3d}: {dt*1e3:8.2f} ms  ({(n-1)/dt:8.0f} pairs/s)  mean pts {cnt.mean():.0f}")
