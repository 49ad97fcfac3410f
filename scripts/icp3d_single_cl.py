"""Experiment: one pair through rst_icp3d_depth against the CTAs per pair (cluster size) of k_icp3d."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from realsensetracker_b200 import Aligner, synth
W, H = 640, 480
intr = synth.intrinsics_for(W, H)
pinned = torch.empty((2, H, W), dtype=torch.int16, pin_memory=True)
frames = pinned.numpy().view(np.uint16)
synth.render_sequence(2, W, H, seed=0, pinned=frames)
al = Aligner(16, 16, 2, 1)
s, d = np.array([1], dtype=np.int32), np.array([0], dtype=np.int32)
for it in (128, 0):
    for c in (1, 2, 4, 8, 16, 0):
        al.set_icp3d_cluster(c)
        al.icp3d_depth(frames, s, d, intr, max_iter=it)
        best = 1e9
        for _ in range(10):
            t0 = time.perf_counter(); al.icp3d_depth(frames, s, d, intr, max_iter=it); best = min(best, time.perf_counter() - t0)
        print(f"iters {it} ctas/pair {c}: {best*1e3:.3f} ms", flush=True)
al.close()
