"""A batch of pairs through rst_icp3d_depth (ncu target for k_icp3d): argv = pairs, iterations, ctas per pair."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from realsensetracker_b200 import Aligner, synth
W, H = 640, 480
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
it = int(sys.argv[2]) if len(sys.argv) > 2 else 32
intr = synth.intrinsics_for(W, H)
frames, gt = synth.render_sequence(n + 1, W, H, seed=0)
al = Aligner(16, 16, 2, 1)
al.set_icp3d_cluster(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
s, d = np.arange(1, n + 1, dtype=np.int32), np.arange(0, n, dtype=np.int32)
for _ in range(2):
    ok, T, mc, cnt = al.icp3d_depth(frames, s, d, intr, max_iter=it)
al.close()
