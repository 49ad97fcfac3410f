"""Experiment: the neighbour cache of k_icp3d (rst_set_icp3d_cache). Batch of 128 pairs and a single pair through
rst_icp3d_depth with the cache off and with several margin settings; checks the poses are bit-identical to cache-off."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from realsensetracker_b200 import Aligner, synth
W, H = 640, 480
intr = synth.intrinsics_for(W, H)
n = 129
import torch
pinned = torch.empty((n, H, W), dtype=torch.int16, pin_memory=True)   # as bench.py holds its frames
frames = pinned.numpy().view(np.uint16)
_, gt = synth.render_sequence(n, W, H, seed=0, pinned=frames)
al = Aligner(16, 16, 2, 1)
s, d = np.arange(1, n, dtype=np.int32), np.arange(0, n - 1, dtype=np.int32)
al.icp3d_depth(frames, s, d, intr, max_iter=1)
settings = [(0, 0, 0), (1, 0.05, 0.2), (4, 0.05, 0.5), (2, 0.1, 0.3), (0, 0.05, 0.05)]
if len(sys.argv) > 1:
    settings = [(0, 0, 0)] + [tuple(float(x) for x in a.split(",")) for a in sys.argv[1:]]
base = None
for cell in (0.1,):
    for st in settings:
        al.set_icp3d_cache(*st)
        best, one = 1e9, 1e9
        for _ in range(3):
            t0 = time.perf_counter(); ok, T, mc, cnt = al.icp3d_depth(frames, s, d, intr, max_iter=128, grid_cell=cell); best = min(best, time.perf_counter() - t0)
        if base is None:
            base = T.copy()
        same = np.array_equal(T, base)
        sq = al.icp3d_cache_stats()
        for _ in range(5):
            t0 = time.perf_counter(); al.icp3d_depth(frames[:2], s[:1], d[:1], intr, max_iter=128, grid_cell=cell); one = min(one, time.perf_counter() - t0)
        print(f"cell {cell} cache {st}: 128 pairs {best*1e3:8.2f} ms = {128/best:8.0f} pairs/s, single pair {one*1e3:6.2f} ms, bit-identical to cache off: {same}, searched {sq[0] / max(sq[1], 1):.4f} of {sq[1]} queries", flush=True)
al.set_icp3d_cache()
for on in (False, True):
    al.set_icp3d_fixed_point_skip(on)
    best, one = 1e9, 1e9
    for _ in range(3):
        t0 = time.perf_counter(); ok, T, mc, cnt = al.icp3d_depth(frames, s, d, intr, max_iter=128); best = min(best, time.perf_counter() - t0)
    it_stats = al.icp3d_iteration_stats()
    for _ in range(5):
        t0 = time.perf_counter(); al.icp3d_depth(frames[:2], s[:1], d[:1], intr, max_iter=128); one = min(one, time.perf_counter() - t0)
    print(f"fixed-point skip {on}: 128 pairs {best*1e3:8.2f} ms = {128/best:8.0f} pairs/s, single pair {one*1e3:6.2f} ms, iterations run {it_stats}, bit-identical: {np.array_equal(T, base)}", flush=True)
for it in (0, 1, 2, 4, 8, 16, 32, 64, 128):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); al.icp3d_depth(frames, s, d, intr, max_iter=it); best = min(best, time.perf_counter() - t0)
    sq = al.icp3d_cache_stats()
    print(f"iters {it:3d}: {best*1e3:8.2f} ms, searched {sq[0]} of {sq[1]}", flush=True)
al.close()
