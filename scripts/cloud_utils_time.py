"""Experiment: wall-clock of the cloud utilities (host clouds in, results out) on one 640x480 frame pair's 5 cm voxel clouds."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from realsensetracker_b200 import Aligner, synth
W, H = 640, 480
intr = synth.intrinsics_for(W, H)
frames, gt = synth.render_sequence(2, W, H, seed=0)
al = Aligner(16, 16, 2, 1)
ok, T, mc, cnt = al.icp3d_depth(frames, [1], [0], intr, max_iter=1)
src, dst = al.icp3d_read_cloud(1, int(cnt[1])), al.icp3d_read_cloud(0, int(cnt[0]))
print("clouds", src.shape, dst.shape)
def tm(name, f, reps=5):
    f(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); best = min(best, time.perf_counter() - t0)
    print(f"{name:32s} {best*1e3:8.3f} ms", flush=True)
    return r
tm("find_correspondences", lambda: al.find_correspondences(dst, src))
tm("cloud_normals k=16", lambda: al.cloud_normals(src, k=16))
cs = tm("cloud_covariances", lambda: al.cloud_covariances(src))
tm("cloud_covariances gicp", lambda: al.cloud_covariances(src, use_gicp=True))
tm("downsample_voxel 0.1", lambda: al.downsample_voxel(src, 0.1))
tm("remove_nans", lambda: al.remove_nans(src))
tm("icp3d_pairs 128 it", lambda: al.icp3d_pairs([src], [dst], 128))
Tg, st = tm("gicp_align 16 x 4", lambda: al.gicp_align(src, dst))
print("gicp pose err vs gt", synth.pose_error(Tg, gt[0]), "icp3d", synth.pose_error(al.icp3d_pairs([src], [dst], 128)[1][0], gt[0]))
al.close()
