import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from realsensetracker_b200 import Aligner, synth
W, H = 640, 480
intr = synth.intrinsics_for(W, H)
frames, gt = synth.render_sequence(17, W, H, seed=0)
al = Aligner(16, 16, 2, 1)
s, d = np.arange(1, 17, dtype=np.int32), np.arange(0, 16, dtype=np.int32)
ok, T, mc, cnt = al.icp3d_depth(frames, s, d, intr, max_iter=16)
print(ok.all(), cnt.mean())
