import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from realsensetracker_b200 import Aligner, synth
W, H = 640, 480
intr = synth.intrinsics_for(W, H)
n = 129
pinned = torch.empty((n, H, W), dtype=torch.int16, pin_memory=True)
frames = pinned.numpy().view(np.uint16)
synth.render_sequence(n, W, H, seed=0, pinned=frames)
al = Aligner(16, 16, 2, 1)
for npairs in (16, 32, 64, 96, 128):
    s, d = np.arange(1, npairs + 1, dtype=np.int32), np.arange(0, npairs, dtype=np.int32)
    al.set_icp3d_cluster(1)
    res = []
    for it in (8, 136):
        al.icp3d_depth(frames[:npairs + 1], s, d, intr, max_iter=it)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); al.icp3d_depth(frames[:npairs + 1], s, d, intr, max_iter=it); best = min(best, time.perf_counter() - t0)
        res.append(best)
    print(f"{npairs} pairs, one CTA each: {(res[1]-res[0])/128*1e6:.1f} us per late iteration", flush=True)
al.close()
