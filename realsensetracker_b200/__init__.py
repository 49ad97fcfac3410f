"""realsensetracker_b200 — B200-native RGB-D frame alignment behind the RealsenseTracker align API.

The compute path is the CUDA library csrc/ -> _lib/librst_align.so (C ABI: include/rst_align.h).
Importing the package does not load it; the first use does, and raises if it is missing.
"""
from . import synth  # noqa: F401
from .align import Aligner, AlignRgbd, AlignIcp3d, default_params, RstError  # noqa: F401
from ._native import (RST_STATUS_OK, RST_STATUS_TOO_FEW, RST_STATUS_DEGENERATE, RST_STATUS_NON_FINITE,  # noqa: F401
                      RST_ROBUST_NONE, RST_ROBUST_HUBER, RST_ROBUST_GEMAN_MCCLURE, Intrinsics, Params, Stats)

__version__ = "0.1.0"
