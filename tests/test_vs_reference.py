"""The headline path (projective point-to-plane pyramid) pinned to the reference: on the SAME depth frames the
reference's own align path (rs_tracker/align/src/align_icp.cpp + common/src/point_cloud_utils.cpp compiled
unmodified -> oracle/_ref/libref.so, run as its caller runs it: back-project -> RemoveNans -> DownsampleVoxel(0.05)
-> AlignIcp3d(128), rs_replay_app.cpp:229,246-251) and this repo's path are both scored against the known motion,
and the product must not be worse than the reference.

The two are different algorithms by the north star's own definition (the reference is KD-tree point-to-point on
5 cm voxels), so "poses within 1e-4 m of the reference" cannot hold literally: the reference itself is 1-2 cm from
the truth. What is asserted: err(product vs truth) <= err(reference vs truth) per pair, for translation and
rotation, and both absolute errors inside stated bounds. CPU: the specification (Oracle-N). GPU: the CUDA path.
"""
import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O
from realsensetracker_b200 import synth

GN = np.load(ROOT / "tests" / "golden" / "oracle_n_160x120.npz")


def reference_poses(src, dst, intr):
    """The compiled reference when oracle/_ref/libref.so is there, else its bit-identical restatement."""
    if O.ref_lib() is not None:
        ok, T = O.ref_align_depth_pairs(src, dst, intr, n_threads=4)
        return ok, T, "reference (oracle/_ref)"
    ok, T = O.align_depth_pairs(src, dst, intr, voxel=0.05, max_iter=128, n_threads=4)
    return ok, T, "Oracle-R (port)"


def check_not_worse(T_ours, T_ref, gt, abs_t, abs_r, slack=1e-4):
    rows = []
    for i in range(len(gt)):
        eo, er = synth.pose_error(T_ours[i], gt[i]), synth.pose_error(T_ref[i], gt[i])
        rows.append((eo, er))
        assert eo[0] <= er[0] + slack and eo[1] <= er[1] + slack, f"pair {i}: ours {eo} worse than the reference {er}"
        assert eo[0] < abs_t and eo[1] < abs_r, f"pair {i}: ours {eo}"
    return rows


def test_specification_not_worse_than_the_reference_160x120():
    f, intr, gt = GN["frames"], tuple(GN["intr"]), GN["gt"]
    ok, T_ref, _ = reference_poses(f[1:3], f[0:2], intr)
    assert ok.all()
    T_n = np.stack([O.align_pair(f[i + 1], f[i], intr, O.default_params())[0] for i in range(2)])
    rows = check_not_worse(T_n, T_ref, gt, 3e-3, 3e-3)
    # the reference really is centimetre-grade here: the pin is meaningful, not vacuous
    assert max(r[1][0] for r in rows) > 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("size", ["160x120", "640x480"])
def test_gpu_not_worse_than_the_compiled_reference_on_the_same_frames(size, seq640):
    from realsensetracker_b200 import Aligner, default_params
    if size == "160x120":
        frames, intr, gt = GN["frames"], tuple(GN["intr"]), GN["gt"]
        bounds = (3e-3, 3e-3)
    else:
        frames, gt, intr = seq640
        bounds = (2e-4, 1e-3)
    n, h, w = frames.shape
    ok, T_ref, kind = reference_poses(frames[1:], frames[:-1], intr)
    assert ok.all()
    al = Aligner(w, h, n, n - 1)
    try:
        T, st = al.align_sequence(frames, intr, default_params())
    finally:
        al.close()
    assert all(s.status == 0 for s in st)
    rows = check_not_worse(T, T_ref, gt, *bounds)
    worst_ours = max(r[0][0] for r in rows), max(r[0][1] for r in rows)
    worst_ref = max(r[1][0] for r in rows), max(r[1][1] for r in rows)
    print(f"{size}: ours {worst_ours}, {kind} {worst_ref}")
    assert worst_ref[0] > worst_ours[0]
