"""Experiment: device time of one 640x480 pair under the launch schedules (per-iteration graph vs fused cluster kernel)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from realsensetracker_b200 import Aligner, default_params, synth
W, H = 640, 480
intr = synth.intrinsics_for(W, H)
pinned = torch.empty((2, H, W), dtype=torch.int16, pin_memory=True)
frames = pinned.numpy().view(np.uint16)
_, gt = synth.render_sequence(2, W, H, seed=0, pinned=frames)
for name, sched, cl in (("per-iteration (graph + PDL)", 2, 0), ("fused, cluster 8", 1, 8), ("fused, cluster 16", 1, 16), ("hybrid, cluster 16", 3, 16)):
    al = Aligner(W, H, 2, 1)
    P = default_params(tiling=1)
    al.set_schedule(sched)
    if cl: al.set_cluster_size(1, cl)
    for _ in range(20): T, st = al.align_sequence(frames, intr, P)
    t0 = time.perf_counter(); n = 200
    for _ in range(n): T, st = al.align_sequence(frames, intr, P)
    dt = (time.perf_counter() - t0) / n
    print(f"{name:30s}: blocking pair {dt*1e6:7.1f} us, err {synth.pose_error(T[0], gt[0])[0]:.2e}")
    al.close()
