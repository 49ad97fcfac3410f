"""One blocking single-pair alignment at 640x480, repeated (profiling target for the latency path)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from realsensetracker_b200 import Aligner, default_params, synth
W, H = 640, 480
intr = synth.intrinsics_for(W, H)
P = default_params(tiling=int(sys.argv[1]) if len(sys.argv) > 1 else 1)
pinned = torch.empty((2, H, W), dtype=torch.int16, pin_memory=True)
frames = pinned.numpy().view(np.uint16)
synth.render_sequence(2, W, H, seed=0, pinned=frames)
al = Aligner(W, H, 2, 1)
al.set_graph_max_pairs(0)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    al.align_sequence(frames, intr, P)
al.close()
