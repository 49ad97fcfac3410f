"""GPU: robustness of the entry points (round-1 advisor findings): non-finite cloud inputs, slot ranges of
device-bound frames, one outstanding asynchronous call per context, and the success criterion of a pair whose
coarse level failed."""
import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O
from realsensetracker_b200 import Aligner, default_params, synth
from realsensetracker_b200 import _native as N
from realsensetracker_b200.align import RstError

pytestmark = pytest.mark.gpu
GOLD = np.load(ROOT / "tests" / "golden" / "oracle_r.npz")


def test_cloud_icp_survives_non_finite_points_and_poses():
    """A NaN/Inf source point or initial pose has no nearest neighbour: the call must report failure (ok = 0), never
    read out of bounds, and leave the context usable (the reference's callers run RemoveNans first,
    rs_replay_app.cpp:246)."""
    src, dst = GOLD["src"].copy(), GOLD["dst"]
    al = Aligner(16, 16, 2, 1)
    try:
        bad = src.copy(); bad[5] = np.nan; bad[17, 1] = np.inf
        ok, T = al.icp3d_pairs([bad], [dst], 8)
        assert not ok[0]
        T0 = np.eye(4); T0[0, 3] = np.nan
        ok, T = al.icp3d_pairs([src], [dst], 8, T0=T0)
        # iteration 0 finds no neighbour for any point (all weights 0 -> zero covariance -> R = I, t = centroid
        # difference), after which the iterations proceed from a finite pose: either outcome is fine, a fault is not
        assert (not ok[0]) or np.isfinite(T[0]).all()
        bad_dst = dst.copy(); bad_dst[3] = np.nan
        ok, T = al.icp3d_pairs([src], [bad_dst], 8)          # a NaN target point is never anybody's neighbour
        assert np.isfinite(T[0]).all()
        ok, T = al.icp3d_pairs([src], [dst], 128)            # the context still works, and still gets the right answer
        ok_o, T_o = O.align_icp3d(src, dst, 128)
        assert ok[0] and ok_o
        dt, dr = synth.pose_error(T[0], T_o)
        assert dt < 1e-4 and dr < 1e-4
    finally:
        al.close()


def test_gridded_cloud_with_infinite_points_or_a_tiny_cell():
    """The uniform grid is laid over the FINITE coordinates of the gridded (target) cloud and its cell count is
    bounded before the conversion to int: an infinite point or a grid_cell far below the cloud's resolution must
    neither overflow the grid dimensions nor change anybody's nearest neighbour (an infinite point is at infinite
    distance from every query)."""
    src, dst = GOLD["src"], GOLD["dst"]
    far = np.array([[np.inf, 0.0, 0.0], [0.0, -np.inf, 0.0]], np.float32)
    dst_bad = np.vstack([dst, far])                           # appended: the indices of the finite points are unchanged
    al = Aligner(16, 16, 2, 1)
    try:
        idx0, d20 = al.find_correspondences(dst, src)
        idx1, d21 = al.find_correspondences(dst_bad, src)
        assert np.array_equal(idx0, idx1) and np.array_equal(d20, d21)
        idx2, d22 = al.find_correspondences(dst, src, grid_cell=1e-12)     # the cell is grown until the grid fits
        assert np.array_equal(idx0, idx2) and np.array_equal(d20, d22)
        idx3, d23 = al.find_correspondences(np.full((4, 3), np.inf, np.float32), src)   # no finite point at all: one cell
        assert ((idx3 >= 0) & (idx3 < 4)).all()
        n0 = al.cloud_normals(dst, 16)
        n1 = al.cloud_normals(dst_bad, 16)
        assert np.allclose(n0, n1[:len(dst)], atol=1e-6)
        c1 = al.cloud_covariances(dst_bad)
        assert np.isfinite(c1[:len(dst)]).all()
        ok0, T0 = al.icp3d_pairs([src], [dst], 32)
        ok1, T1 = al.icp3d_pairs([src], [dst_bad], 32)
        assert ok0[0] and ok1[0]
        dt, dr = synth.pose_error(T1[0], T0[0])
        assert dt < 1e-5 and dr < 1e-5
    finally:
        al.close()


def test_device_bound_frames_reject_slots_outside_the_bound_range(seq_small):
    import torch
    frames, gt, intr = seq_small
    n, h, w = frames.shape
    pitch = (w + 7) // 8 * 8
    buf = torch.zeros((2, h, pitch), dtype=torch.int16, device="cuda")
    buf[:, :, :w] = torch.from_numpy(frames[:2].view(np.int16)).cuda()
    al = Aligner(w, h, 8, 4)
    try:
        P = default_params()
        al.begin(w, h, intr, P)
        al.set_frames_device(buf.data_ptr(), 2, pitch, pitch * h, first_slot=2)      # slots 2 and 3 only
        al.preprocess(2, 2)
        T, st = al.align_slots([3], [2])
        assert st[0].status == 0
        with pytest.raises(RstError):
            al.preprocess(0, 2)                                                      # slots 0, 1 are not backed by anything
        with pytest.raises(RstError):
            al.align_slots([3], [1])
        with pytest.raises(RstError):
            al.evaluate(4, 2, 0, np.eye(4))
    finally:
        al.close()


def test_one_outstanding_async_call_per_context(seq_small):
    frames, gt, intr = seq_small
    n, h, w = frames.shape
    al = Aligner(w, h, n, n - 1)
    try:
        P = default_params()
        al.submit_sequence(frames, intr, P)
        with pytest.raises(RstError):
            al.submit_sequence(frames, intr, P)          # would overwrite staging an enqueued copy still uses
        with pytest.raises(RstError):
            al.align_sequence(frames, intr, P)
        T, st = al.wait()
        T2, _ = al.align_sequence(frames, intr, P)       # after rst_wait the context is free again
        assert np.array_equal(T, T2)
    finally:
        al.close()


def test_failed_coarse_level_does_not_fail_a_pair_that_converged():
    """Sparse depth: the coarsest level has fewer associations than min_count (TOO_FEW), the finer levels converge.
    status (the success criterion) is the LAST evaluated iteration; the failure stays visible in any_status /
    failed_iterations. Same semantics as the CPU specification."""
    w, h = 160, 120
    intr = (96.0, 96.0, 80.0, 60.0)
    sc = synth.Scene(7)
    Twc = synth.trajectory(2, seed=7, step_t=0.01, step_r=0.008)
    frames = np.stack([sc.render(Twc[k], w, h, intr=intr) for k in range(2)])
    gt = synth.relative_pose(Twc[0], Twc[1])
    kw = dict(min_count=900)                                      # level 2 is 40x30 = 1200 px: fewer than 900 carry a normal
    P, Po = default_params(**kw), O.default_params(**kw)
    To, so = O.align_pair(frames[1], frames[0], intr, Po)
    assert so.status == 0 and so.any_status == N.RST_STATUS_TOO_FEW and so.failed_iterations == Po.iters[2]
    al = Aligner(w, h, 2, 1)
    try:
        T, st = al.align_sequence(frames, intr, P)
        assert st[0].status == 0 and st[0].any_status == N.RST_STATUS_TOO_FEW
        assert st[0].failed_iterations == so.failed_iterations and st[0].iterations == so.iterations
        dt, dr = synth.pose_error(T[0], To)
        assert dt < 1e-4 and dr < 1e-4
        assert synth.pose_error(T[0], gt)[0] < 3e-3
    finally:
        al.close()


def test_graph_replay_is_bit_identical_to_direct_launches(seq_small):
    """Small blocking calls replay a CUDA graph of their kernels (the latency path). Same kernels, same arguments:
    the results must be bit-identical to the direct launches, call after call, for sequences and pairs, and a
    parameter change must re-capture rather than replay a stale graph."""
    frames, gt, intr = seq_small
    n, h, w = frames.shape
    al = Aligner(w, h, 2 * n, n)
    try:
        for tiling in (0, 1):
            P = default_params(tiling=tiling)
            al.set_graph_max_pairs(0)
            T_d, st_d = al.align_sequence(frames, intr, P)
            Tp_d, _ = al.align_pairs(frames[1:3], frames[0:2], intr, P)
            l0 = al.launch_count
            al.align_sequence(frames, intr, P)
            direct_launches = al.launch_count - l0
            al.set_graph_max_pairs(8)
            for _ in range(3):                                   # capture, then two replays
                T_g, st_g = al.align_sequence(frames, intr, P)
                assert np.array_equal(T_g, T_d)
                assert [s.count for s in st_g] == [s.count for s in st_d] and all(s.status == 0 for s in st_g)
            l0 = al.launch_count
            al.align_sequence(frames, intr, P)
            assert al.launch_count - l0 == direct_launches       # a replay accounts for the launches it contains
            Tp_g, _ = al.align_pairs(frames[1:3], frames[0:2], intr, P)
            assert np.array_equal(Tp_g, Tp_d)
        P2 = default_params(iters=[3, 2, 1, 0])
        al.set_graph_max_pairs(0)
        T2_d, _ = al.align_sequence(frames, intr, P2)
        al.set_graph_max_pairs(8)
        T2_g, _ = al.align_sequence(frames, intr, P2)
        assert np.array_equal(T2_g, T2_d) and not np.array_equal(T2_g, T_d)
        T0 = np.stack([gt[i] for i in range(n - 1)])              # the initial pose is data, not part of the graph
        al.set_graph_max_pairs(0)
        T3_d, _ = al.align_sequence(frames, intr, P2, T0=T0)
        al.set_graph_max_pairs(8)
        T3_g, _ = al.align_sequence(frames, intr, P2, T0=T0)
        assert np.array_equal(T3_g, T3_d) and not np.array_equal(T3_g, T2_g)
    finally:
        al.close()


def test_frames_bound_in_place_with_ragged_width_and_dirty_padding():
    """Level 0 read in place from caller memory whose rows are wider than the image, with plausible depth values in the
    padding, a width that is not a multiple of 8 and a first slot > 0: the TMA tensor of k_preprocess (image width as
    its extent) and the bulk-staged depth tile of k_icp_iter (padding masked after the copy) must both treat the
    padding as invalid — geometry maps, association and poses bit-identical to the upload path."""
    import torch
    from realsensetracker_b200 import synth
    w, h = 204, 150                                    # 204 % 8 == 4
    intr = (122.0, 122.0, 101.5, 74.5)
    scene = synth.Scene(5)
    Twc = synth.trajectory(3, seed=5, step_t=0.02, step_r=0.015)
    frames = np.stack([scene.render(Twc[k], w, h, intr=intr) for k in range(3)])
    P = default_params()
    al = Aligner(w, h, 4, 3)
    try:
        T_up, st_up = al.align_sequence(frames, intr, P)
        geo_up = [al.read_geometry(0, l).copy() for l in range(3)]
        idx_up, ev_up = al.evaluate(1, 0, 0, np.eye(4))
        pitch = 224                                    # > roundup(w, 8) = 208
        buf = torch.full((3, h, pitch), 2500, dtype=torch.int16, device="cuda")   # 2.5 m everywhere, padding included
        buf[:, :, :w] = torch.from_numpy(frames.view(np.int16)).cuda()
        al.begin(w, h, intr, P)
        al.set_frames_device(buf.data_ptr(), 3, pitch, pitch * h, first_slot=1)    # slots 1, 2, 3
        al.preprocess(1, 3)
        for l in range(3):
            assert np.array_equal(al.read_geometry(1, l).view(np.uint32), geo_up[l].view(np.uint32)), f"geometry level {l}"
        idx_d, ev_d = al.evaluate(2, 1, 0, np.eye(4))
        assert np.array_equal(idx_d, idx_up) and ev_d.count == ev_up.count
        T_d, st_d = al.align_slots([2, 3], [1, 2])
        assert np.array_equal(T_d, T_up)
        assert [s.count for s in st_d] == [s.count for s in st_up]
    finally:
        al.close()


def test_two_devices_in_one_process(seq_small):
    """One process, one context per GPU: kernel attributes (dynamic shared memory of the photometric variant, opt-in
    cluster sizes) are per device, and both devices must give the same bits."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    frames, gt, intr = seq_small
    n, h, w = frames.shape
    rgb = np.ascontiguousarray(np.repeat((frames >> 4).astype(np.uint8)[..., None], 3, axis=-1))
    src, dst = O_clouds()
    out = []
    for dev in (0, 1):
        al = Aligner(w, h, 2 * n, n, device=dev)
        try:
            P = default_params()
            T, _ = al.align_sequence(frames, intr, P)
            Pp = default_params(photo_weight=0.5)
            Tp, _ = al.align_pairs(frames[1:], frames[:-1], intr, Pp, src_rgb=rgb[1:], dst_rgb=rgb[:-1])
            al.set_icp3d_cluster(16)
            ok, Tc = al.icp3d_pairs([src], [dst], 16)
            out.append((T, Tp, Tc))
        finally:
            al.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


def O_clouds():
    from conftest import ROOT
    g = np.load(ROOT / "tests" / "golden" / "oracle_r.npz")
    return g["src"], g["dst"]


@pytest.mark.parametrize("size", [(72, 40), (33, 17), (130, 66)])
def test_small_and_odd_frames_match_the_oracle(size):
    """Frames smaller than one pre-processing tile / TMA box (80 x 34) and odd sizes: pyramid, geometry maps and
    association bit-exact against Oracle-N on every level, poses within 1e-4."""
    from oracle import oracle as O
    from realsensetracker_b200 import synth
    w, h = size
    levels = 3 if min(w, h) >= 32 else 2
    intr = (0.6 * w, 0.6 * w, 0.5 * w - 0.5, 0.5 * h - 0.5)
    scene = synth.Scene(7)
    Twc = synth.trajectory(2, seed=7, step_t=0.01, step_r=0.01)
    frames = np.stack([scene.render(Twc[k], w, h, intr=intr) for k in range(2)])
    iters = [4, 3, 2, 0][:levels] + [0] * (4 - levels)
    P, Po = default_params(num_levels=levels, iters=iters, min_count=6), O.default_params(num_levels=levels, iters=iters, min_count=6)
    al = Aligner(w, h, 2, 1)
    try:
        al.begin(w, h, intr, P)
        al.upload(frames)
        al.preprocess(0, 2)
        d = frames[0]
        for l in range(levels):
            L = O.level_info(intr, w, h, l)
            if l > 0:
                d = O.pyr_down(d, Po.pyr_depth_tol)
            assert np.array_equal(al.read_depth(0, l), d)
            assert np.array_equal(al.read_geometry(0, l).view(np.uint32), O.geometry(d, L, Po).view(np.uint32)), f"level {l}"
        L0 = O.level_info(intr, w, h, 0)
        idx_o, st_o = O.evaluate(frames[1], None, O.geometry(frames[0], L0, Po), L0, Po, np.eye(4))
        idx_g, st_g = al.evaluate(1, 0, 0, np.eye(4))
        assert np.array_equal(idx_g, idx_o) and st_g.count == st_o.count
        T, st = al.align_sequence(frames, intr, P)
        To, so = O.align_pair(frames[1], frames[0], intr, Po)
        if so.status == 0:
            dt, dr = synth.pose_error(T[0], To)
            assert st[0].status == 0 and dt < 1e-4 and dr < 1e-4, (dt, dr)
        else:
            assert st[0].status != 0
    finally:
        al.close()


def test_blocking_call_on_a_large_batch_pipelines_its_upload_without_changing_bits():
    """A blocking host-frame call on >= 64 frames / 32 pairs runs as two halves (upload of the second under the kernels of
    the first, rst_set_pipeline_chunk automatic): same bits as the unchunked call and as the asynchronous form."""
    from realsensetracker_b200 import synth
    w, h, n = 208, 152, 70
    intr = synth.intrinsics_for(w, h)
    frames, gt = synth.render_sequence(n, w, h, seed=4, step_t=0.01, step_r=0.008)
    P = default_params()
    al = Aligner(w, h, 2 * n, n)
    try:
        al.set_pipeline_chunk(-1)
        T_plain, st_plain = al.align_sequence(frames, intr, P)
        Tp_plain, _ = al.align_pairs(frames[1:], frames[:-1], intr, P)
        al.set_pipeline_chunk(0)                          # automatic (the default)
        l0 = al.launch_count
        T_auto, st_auto = al.align_sequence(frames, intr, P)
        assert al.launch_count - l0 > 30                  # two halves: more launches than the 23 of one pass
        Tp_auto, _ = al.align_pairs(frames[1:], frames[:-1], intr, P)
        assert np.array_equal(T_auto, T_plain) and np.array_equal(Tp_auto, Tp_plain) and np.array_equal(Tp_plain, T_plain)
        assert [s.count for s in st_auto] == [s.count for s in st_plain]
        al.submit_sequence(frames, intr, P)
        T_async, _ = al.wait()
        assert np.array_equal(T_async, T_plain)
        # the same with colour frames and the photometric term (RGB is uploaded chunk by chunk too)
        rgb = np.ascontiguousarray(np.repeat((frames >> 4).astype(np.uint8)[..., None], 3, axis=-1))
        Pp = default_params(photo_weight=0.5)
        al.set_pipeline_chunk(-1)
        Tc_plain, _ = al.align_pairs(frames[1:41], frames[0:40], intr, Pp, src_rgb=rgb[1:41], dst_rgb=rgb[0:40])
        al.set_pipeline_chunk(0)
        Tc_auto, _ = al.align_pairs(frames[1:41], frames[0:40], intr, Pp, src_rgb=rgb[1:41], dst_rgb=rgb[0:40])
        assert np.array_equal(Tc_auto, Tc_plain) and not np.array_equal(Tc_plain, Tp_plain[:40])
    finally:
        al.close()
