"""Experiment: latency of one blocking AlignRgbd call (the reference's online use: one pair at a time)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from realsensetracker_b200 import Aligner, default_params, synth

for (W, H) in ((640, 480), (1280, 720)):
    intr = synth.intrinsics_for(W, H)
    P = default_params(tiling=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    pinned = torch.empty((2, H, W), dtype=torch.int16, pin_memory=True)
    frames = pinned.numpy().view(np.uint16)
    _, gt = synth.render_sequence(2, W, H, seed=0, pinned=frames)
    al = Aligner(W, H, 2, 1)
    for _ in range(20): T, st = al.align_sequence(frames, intr, P)
    t0 = time.perf_counter(); n = 200
    for _ in range(n): T, st = al.align_sequence(frames, intr, P)
    dt = (time.perf_counter() - t0) / n
    print(f"{W}x{H}: blocking single-pair call {dt*1e6:.1f} us  ({1/dt:.0f} pairs/s serial), err {synth.pose_error(T[0], gt[0])}")
    al.begin(W, H, intr, P); al.upload(frames); al.sync()
    def comp():
        al.preprocess(0, 2); al.align_slots([1], [0], fetch=False); al.sync()
    for _ in range(20): comp()
    t0 = time.perf_counter()
    for _ in range(n): comp()
    print(f"   device-resident compute only {(time.perf_counter()-t0)/n*1e6:.1f} us")
    al.close()
