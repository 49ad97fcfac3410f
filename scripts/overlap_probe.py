"""Experiment: where does the e2e time go (H2D alone, compute alone, pipelined)?"""
import sys, time, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from realsensetracker_b200 import Aligner, default_params, synth
from realsensetracker_b200 import _native as N
from realsensetracker_b200.align import _frames

W, H, F = 640, 480, 129
intr = synth.intrinsics_for(W, H)
P = default_params()
pinned = torch.empty((F, H, W), dtype=torch.int16, pin_memory=True)
frames = pinned.numpy().view(np.uint16)
synth.render_sequence(F, W, H, seed=0, pinned=frames)
stream = torch.cuda.Stream()
al = Aligner(W, H, F, F - 1, stream=stream.cuda_stream)
src, dst = np.arange(1, F, dtype=np.int32), np.arange(0, F - 1, dtype=np.int32)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def up_only():
    al.begin(W, H, intr, P); al.upload(frames); al.sync()
def comp_only():
    al.preprocess(0, F); al.align_slots(src, dst, fetch=False); al.sync()
def up_comp():
    al.begin(W, H, intr, P); al.upload(frames); al.preprocess(0, F); al.align_slots(src, dst)
print("upload only ms", timeit(up_only))
print("compute only ms", timeit(comp_only))
print("upload+compute+fetch serial ms", timeit(up_comp))
d = torch.empty((F, H, W), dtype=torch.int16, device="cuda")
def torch_copy():
    d.copy_(pinned, non_blocking=True); torch.cuda.synchronize()
print("torch H2D ms", timeit(torch_copy), "GB/s", F*H*W*2/1e6/timeit(torch_copy))
for ch in (0, 64, 32):
    al.set_pipeline_chunk(ch)
    print("align_sequence chunk", ch, "ms", timeit(lambda: al.align_sequence(frames, intr, P)))
# half-size compute: is a 64-pair launch half the time of a 128-pair launch?
def comp_half():
    al.preprocess(0, 65); al.align_slots(src[:64], dst[:64], fetch=False); al.sync()
print("compute 64 pairs ms", timeit(comp_half))
def comp_q():
    al.preprocess(0, 33); al.align_slots(src[:32], dst[:32], fetch=False); al.sync()
print("compute 32 pairs ms", timeit(comp_q))
