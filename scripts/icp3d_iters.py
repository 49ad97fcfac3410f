"""Experiment: rst_icp3d_depth time against the iteration count (fixed cost vs per-iteration cost)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from realsensetracker_b200 import Aligner, synth
W, H = 640, 480
intr = synth.intrinsics_for(W, H)
n = 129
frames, gt = synth.render_sequence(n, W, H, seed=0)
al = Aligner(16, 16, 2, 1)
s, d = np.arange(1, n, dtype=np.int32), np.arange(0, n - 1, dtype=np.int32)
al.icp3d_depth(frames, s, d, intr, max_iter=1)
for it in (0, 1, 2, 8, 32, 128):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); al.icp3d_depth(frames, s, d, intr, max_iter=it); best = min(best, time.perf_counter() - t0)
    print(f"iters {it:3d}: {best*1e3:8.2f} ms")
