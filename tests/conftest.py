import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build what is missing (a no-op when the .so files travelled with the snapshot)."""
    from realsensetracker_b200 import build
    build.build_synth()
    if not (build.LIBDIR / "librst_align.so").exists():
        build.build_align()
    from oracle import oracle as O
    O.build()


@pytest.fixture(scope="session")
def seq640():
    """5 frames 640x480 of a smooth trajectory + ground-truth frame-to-frame poses."""
    from realsensetracker_b200 import synth
    frames, gt = synth.render_sequence(5, 640, 480, seed=0)
    return frames, gt, synth.intrinsics_for(640, 480)


@pytest.fixture(scope="session")
def seq_small():
    """4 frames 208x152 (not a multiple of 64 or 32: exercises tile and chunk tails)."""
    from realsensetracker_b200 import synth
    w, h = 208, 152
    intr = (125.0, 125.0, 104.0, 76.0)
    scene = synth.Scene(3)
    Twc = synth.trajectory(4, seed=3, step_t=0.02, step_r=0.015)
    frames = np.stack([scene.render(Twc[k], w, h, intr=intr) for k in range(4)])
    gt = np.stack([synth.relative_pose(Twc[k], Twc[k + 1]) for k in range(3)])
    return frames, gt, intr


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
