"""One pair through rst_icp3d_depth (depth frames -> clouds -> AlignIcp3d), repeated: ncu launch-list target."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from realsensetracker_b200 import Aligner, synth
W, H = 640, 480
intr = synth.intrinsics_for(W, H)
frames, gt = synth.render_sequence(2, W, H, seed=0)
al = Aligner(16, 16, 2, 1)
s, d = np.array([1], dtype=np.int32), np.array([0], dtype=np.int32)
for _ in range(3):
    ok, T, mc, cnt = al.icp3d_depth(frames, s, d, intr)
cloud = al.icp3d_read_cloud(0, int(cnt[0]))
if cloud is not None:
    for _ in range(2):
        al.cloud_normals(cloud, 16)
        al.cloud_covariances(cloud)
        al.find_correspondences(cloud, cloud)
al.close()
