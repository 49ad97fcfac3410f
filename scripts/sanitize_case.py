"""Smallest end-to-end case for compute-sanitizer: one 208x152 sequence (tile/chunk tails), photometric pair, cloud ICP."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from realsensetracker_b200 import Aligner, default_params, synth
from realsensetracker_b200 import _native as N

w, h, intr = 208, 152, (125.0, 125.0, 104.0, 76.0)
sc = synth.Scene(3)
Twc = synth.trajectory(3, seed=3, step_t=0.02, step_r=0.015)
fr = [sc.render(Twc[k], w, h, intr=intr, rgb=True) for k in range(3)]
depth = np.stack([f[0] for f in fr]); rgb = np.stack([f[1] for f in fr])
al = Aligner(w, h, 6, 3)
T, st = al.align_sequence(depth, intr, default_params())
print("seq", [s.status for s in st])
T, st = al.align_pairs(depth[1:], depth[:-1], intr, default_params(robust_kind=N.RST_ROBUST_HUBER, robust_scale=0.01, normal_cos_min=0.8, tiling=1))
print("pairs huber+ngate+latency", [s.status for s in st])
T, st = al.align_pairs(depth[1:], depth[:-1], intr, default_params(photo_weight=0.5), src_rgb=rgb[1:], dst_rgb=rgb[:-1])
print("photo", [s.status for s in st])
idx, s1 = al.evaluate(1, 0, 0, np.eye(4))
print("evaluate", s1.count, int((idx >= 0).sum()))
ok, T, mc, cnt = al.icp3d_depth(depth, [1, 2], [0, 1], intr, max_iter=6)
print("icp3d", ok, cnt)
al.close()
