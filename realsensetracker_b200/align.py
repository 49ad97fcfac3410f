"""Host-side mirror of the align interface over the C ABI (include/rst_align.h).

Python here is only the test/bench harness language; the reference-facing host API is the
C++ header include/rs_tracker/align/align_rgbd.hpp, which calls the same C ABI. Names and
argument meaning follow the reference (`AlignIcp3d(src, dst, max_iter, &transform)`,
align_icp.hpp:19-24): src -> dst pose, initial guess in, result out, bool success.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from ._native import Intrinsics, Params, Stats, Frame, RstError


def default_params(**kw) -> Params:
    p = Params()
    N.align_lib().rst_params_default(C.byref(p))
    for k, v in kw.items():
        if k == "iters":
            for i in range(N.RST_MAX_LEVELS):
                p.iters[i] = v[i] if i < len(v) else 0
        else:
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
    return p


def pose_to_cm(T) -> np.ndarray:
    """[..,4,4] -> [..,16] column-major fp32 (bit-compatible with Eigen::Isometry3f::matrix())."""
    T = np.asarray(T, dtype=np.float32)
    return np.ascontiguousarray(np.swapaxes(T, -1, -2)).reshape(T.shape[:-2] + (16,))


def cm_to_pose(p) -> np.ndarray:
    p = np.asarray(p)
    return np.swapaxes(p.reshape(p.shape[:-1] + (4, 4)), -1, -2).astype(np.float64)


def _frames(arr: np.ndarray, rgb: np.ndarray | None = None):
    """numpy depth [n,h,w] uint16 (any row stride) and optional rgb [n,h,w,3] uint8 -> ctypes Frame array."""
    assert arr.dtype == np.uint16 and arr.ndim == 3 and arr.strides[2] == 2
    n, h, w = arr.shape
    if rgb is not None:
        assert rgb.dtype == np.uint8 and rgb.shape == (n, h, w, 3) and rgb.strides[3] == 1 and rgb.strides[2] == 3
    fr = (Frame * n)()
    base = arr.ctypes.data
    for i in range(n):
        fr[i].depth = base + i * arr.strides[0]
        fr[i].rgb = rgb.ctypes.data + i * rgb.strides[0] if rgb is not None else None
        fr[i].width, fr[i].height = w, h
        fr[i].depth_stride_bytes = arr.strides[1]
        fr[i].rgb_stride_bytes = rgb.strides[1] if rgb is not None else 0
    return fr


def stats_to_dict(s: Stats) -> dict:
    return dict(status=s.status, iterations=s.iterations, count=s.count, rmse=s.rmse, sum_wr2=s.sum_wr2,
                A=np.array(s.A[:]), b=np.array(s.b[:]))


class Aligner:
    """One alignment context on one GPU (rst_ctx). Not thread-safe; contexts are independent."""

    def __init__(self, max_w: int, max_h: int, max_frames: int, max_pairs: int, device: int = 0, stream: int | None = None):
        self._lib = N.align_lib()
        self._ctx = C.c_void_p()
        rc = self._lib.rst_ctx_create(device, max_w, max_h, max_frames, max_pairs, stream, C.byref(self._ctx))
        if rc != N.RST_OK:
            raise RstError(rc, self._lib.rst_last_create_error().decode())
        self.max_frames, self.max_pairs = max_frames, max_pairs

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.rst_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != N.RST_OK:
            raise RstError(rc, self._lib.rst_last_error(self._ctx).decode())

    # ---- one-call API (host frames in, poses out) -------------------------------------------
    def align_pairs(self, src: np.ndarray, dst: np.ndarray, intr, params: Params | None = None, T0=None,
                    src_rgb: np.ndarray | None = None, dst_rgb: np.ndarray | None = None):
        """src, dst: [n,h,w] uint16 (+ optional rgb [n,h,w,3] uint8 for the photometric term).
        Returns (poses [n,4,4], list of Stats)."""
        n = src.shape[0]
        P = params if params is not None else default_params()
        K = Intrinsics(*intr)
        poses = pose_to_cm(np.broadcast_to(np.eye(4) if T0 is None else T0, (n, 4, 4))).copy()
        stats = (Stats * n)()
        self._check(self._lib.rst_align_pairs(self._ctx, _frames(src, src_rgb), _frames(dst, dst_rgb), n, C.byref(K), C.byref(P),
                                              poses.ctypes.data, C.addressof(stats)))
        return cm_to_pose(poses), list(stats)

    def align_sequence(self, frames: np.ndarray, intr, params: Params | None = None, T0=None):
        """frames: [n,h,w] uint16; pair i aligns frames[i+1] onto frames[i]. Returns ([n-1,4,4], stats)."""
        n = frames.shape[0]
        P = params if params is not None else default_params()
        K = Intrinsics(*intr)
        poses = pose_to_cm(np.broadcast_to(np.eye(4) if T0 is None else T0, (n - 1, 4, 4))).copy()
        stats = (Stats * (n - 1))()
        self._check(self._lib.rst_align_sequence(self._ctx, _frames(frames), n, C.byref(K), C.byref(P),
                                                 poses.ctypes.data, C.addressof(stats)))
        return cm_to_pose(poses), list(stats)

    # ---- asynchronous form: submit now, wait later (two contexts ping-pong to overlap PCIe and kernels)
    def submit_sequence(self, frames: np.ndarray, intr, params: Params | None = None):
        """Enqueues upload + alignment + result download of a sequence and returns at once.
        `frames` must stay alive (and unchanged) until wait() returns."""
        P = params if params is not None else default_params()
        K = Intrinsics(*intr)
        self._pending = (frames.shape[0] - 1, _frames(frames), frames)
        self._check(self._lib.rst_align_sequence_async(self._ctx, self._pending[1], frames.shape[0], C.byref(K), C.byref(P), None))

    def wait(self):
        """Blocks until the submitted work is done; returns (poses [n,4,4], list of Stats)."""
        n = self._pending[0]
        poses = np.empty((n, 16), dtype=np.float32)
        stats = (Stats * n)()
        self._check(self._lib.rst_wait(self._ctx, poses.ctypes.data, C.addressof(stats)))
        self._pending = None
        return cm_to_pose(poses), list(stats)

    # ---- cloud-based alignment with the reference's own algorithm (AlignIcp3d, align_icp.cpp:73-167)
    def icp3d_pairs(self, src_clouds, dst_clouds, max_iter: int = 128, T0=None, grid_cell: float = 0.1, details: bool = False):
        """src_clouds/dst_clouds: lists of [n,3] float32 arrays. Returns (ok [n] bool, poses [n,4,4])
        and, with details=True, a list of dicts (mean_cost, cov, nbrs, weights of the last iteration)."""
        n = len(src_clouds)
        S = [np.ascontiguousarray(a, dtype=np.float32) for a in src_clouds]
        D = [np.ascontiguousarray(a, dtype=np.float32) for a in dst_clouds]
        cs, cd = (N.Cloud * n)(), (N.Cloud * n)()
        for i in range(n):
            cs[i].xyz, cs[i].n = S[i].ctypes.data, len(S[i])
            cd[i].xyz, cd[i].n = D[i].ctypes.data, len(D[i])
        poses = pose_to_cm(np.broadcast_to(np.eye(4) if T0 is None else T0, (n, 4, 4))).copy()
        res = (N.Icp3dResult * n)()
        tot = sum(len(a) for a in S)
        nbrs = np.empty(tot, dtype=np.int32) if details else None
        wts = np.empty(tot, dtype=np.float32) if details else None
        self._check(self._lib.rst_icp3d_pairs(self._ctx, cs, cd, n, max_iter, grid_cell, poses.ctypes.data, C.addressof(res),
                                              nbrs.ctypes.data if details else None, wts.ctypes.data if details else None))
        ok = np.array([r.ok != 0 for r in res])
        T = cm_to_pose(poses)
        if not details:
            return ok, T
        out, o = [], 0
        for i in range(n):
            m = len(S[i])
            out.append(dict(mean_cost=res[i].mean_cost, mu=res[i].mu, cov=np.array(res[i].cov[:]).reshape(3, 3),
                            nbrs=nbrs[o:o + m].copy(), weights=wts[o:o + m].copy()))
            o += m
        return ok, T, out

    def solve_kabsch(self, src, dst, pairs, weights=None):
        """SolveKabsch(src, dst, indices, weights, &xfm) (align_icp.cpp:18-71) on the GPU. Returns (ok, pose 4x4)."""
        S, D = np.ascontiguousarray(src, dtype=np.float32), np.ascontiguousarray(dst, dtype=np.float32)
        pr = np.ascontiguousarray(pairs, dtype=np.int32)
        w = np.ascontiguousarray(weights, dtype=np.float32) if weights is not None else None
        cs, cd = N.Cloud(S.ctypes.data, len(S)), N.Cloud(D.ctypes.data, len(D))
        pose = np.zeros(16, dtype=np.float32)
        ok = C.c_int32(0)
        self._check(self._lib.rst_solve_kabsch(self._ctx, C.byref(cs), C.byref(cd), pr.ctypes.data, len(pr),
                                               w.ctypes.data if w is not None else None, pose.ctypes.data, C.byref(ok)))
        return bool(ok.value), cm_to_pose(pose)

    def cloud_normals(self, cloud, k: int = 16, viewpoint=(0.0, 0.0, 0.0), grid_cell: float = 0.0) -> np.ndarray:
        """ComputeNormals + OrientNormals (point_cloud_utils.cpp:176-216) on the GPU: [n,3] float32 normals."""
        S = np.ascontiguousarray(cloud, dtype=np.float32)
        vp = np.ascontiguousarray(viewpoint, dtype=np.float32)
        out = np.empty_like(S)
        cs = N.Cloud(S.ctypes.data, len(S))
        self._check(self._lib.rst_cloud_normals(self._ctx, C.byref(cs), k, vp.ctypes.data, grid_cell, out.ctypes.data))
        return out

    def find_correspondences(self, target, source, grid_cell: float = 0.0):
        """FindCorrespondences(tree(target), source, &indices, &squared_distances) (point_cloud_utils.cpp:70-90) on the GPU."""
        T_, S = np.ascontiguousarray(target, dtype=np.float32), np.ascontiguousarray(source, dtype=np.float32)
        idx = np.empty(len(S), dtype=np.int32); d2 = np.empty(len(S), dtype=np.float32)
        ct, cs = N.Cloud(T_.ctypes.data, len(T_)), N.Cloud(S.ctypes.data, len(S))
        self._check(self._lib.rst_find_correspondences(self._ctx, C.byref(ct), C.byref(cs), grid_cell, idx.ctypes.data, d2.ctypes.data))
        return idx, d2

    def cloud_centroid(self, cloud) -> np.ndarray:
        """ComputeCentroid(cloud, &centroid) (point_cloud_utils.cpp:92-98) on the GPU: [3] float32."""
        S = np.ascontiguousarray(cloud, dtype=np.float32)
        out = np.empty(3, dtype=np.float32)
        cs = N.Cloud(S.ctypes.data, len(S))
        self._check(self._lib.rst_cloud_centroid(self._ctx, C.byref(cs), out.ctypes.data))
        return out

    def tree_create(self, cloud, grid_cell: float = 0.0) -> int:
        """KDTree3f{cloud} (kdtree.hpp:26-35) on the GPU: an opaque handle of the cloud's device-resident search grid."""
        S = np.ascontiguousarray(cloud, dtype=np.float32)
        cs = N.Cloud(S.ctypes.data, len(S))
        h = C.c_void_p()
        self._check(self._lib.rst_tree_create(self._ctx, C.byref(cs), grid_cell, C.byref(h)))
        return h.value

    def tree_query(self, tree: int, queries, k: int = 1):
        """KDTree3f::query (kdtree.hpp:51-57) for a batch: (indices [n,k] int32, squared distances [n,k] float32)."""
        Q = np.ascontiguousarray(queries, dtype=np.float32)
        idx = np.empty((len(Q), k), dtype=np.int32); d2 = np.empty((len(Q), k), dtype=np.float32)
        self._check(self._lib.rst_tree_query(self._ctx, tree, Q.ctypes.data, len(Q), k, idx.ctypes.data, d2.ctypes.data))
        return idx, d2

    def tree_destroy(self, tree: int) -> None:
        self._lib.rst_tree_destroy(tree)

    def cloud_extents(self, cloud):
        """ComputeExtents(cloud, &box) (point_cloud_utils.cpp:26-32) on the GPU: (lo [3], hi [3]) float32."""
        S = np.ascontiguousarray(cloud, dtype=np.float32)
        lo, hi = np.empty(3, dtype=np.float32), np.empty(3, dtype=np.float32)
        cs = N.Cloud(S.ctypes.data, len(S))
        self._check(self._lib.rst_cloud_extents(self._ctx, C.byref(cs), lo.ctypes.data, hi.ctypes.data))
        return lo, hi

    def orient_normals(self, cloud, viewpoint, normals) -> np.ndarray:
        """OrientNormals(cloud, viewpoint, &normals) (point_cloud_utils.cpp:205-216) on the GPU: a flipped copy of `normals`."""
        S = np.ascontiguousarray(cloud, dtype=np.float32)
        vp = np.ascontiguousarray(viewpoint, dtype=np.float32)
        out = np.array(normals, dtype=np.float32, order="C", copy=True)
        if out.shape != S.shape:
            raise ValueError("normals must have the cloud's shape")
        cs = N.Cloud(S.ctypes.data, len(S))
        self._check(self._lib.rst_orient_normals(self._ctx, C.byref(cs), vp.ctypes.data, out.ctypes.data))
        return out

    def cloud_covariances(self, cloud, use_gicp: bool = False, grid_cell: float = 0.0) -> np.ndarray:
        """ComputeCovariances(tree, cloud, &covs, use_gicp) (point_cloud_utils.cpp:100-161) on the GPU: [n,3,3] float32."""
        S = np.ascontiguousarray(cloud, dtype=np.float32)
        out = np.empty((len(S), 3, 3), dtype=np.float32)
        cs = N.Cloud(S.ctypes.data, len(S))
        self._check(self._lib.rst_cloud_covariances(self._ctx, C.byref(cs), 1 if use_gicp else 0, grid_cell, out.ctypes.data))
        return out

    def downsample_voxel(self, cloud, voxel: float) -> np.ndarray:
        """DownsampleVoxel (point_cloud_utils.cpp:34-68) on the GPU; first point per voxel, first-occurrence order."""
        S = np.ascontiguousarray(cloud, dtype=np.float32)
        out = np.empty_like(S); n = C.c_int32(0)
        cs = N.Cloud(S.ctypes.data, len(S))
        self._check(self._lib.rst_downsample_voxel(self._ctx, C.byref(cs), voxel, out.ctypes.data, C.byref(n)))
        return out[:n.value].copy()

    def remove_nans(self, cloud) -> np.ndarray:
        """RemoveNans (point_cloud_utils.cpp:163-174) on the GPU."""
        S = np.ascontiguousarray(cloud, dtype=np.float32)
        out = np.empty_like(S); n = C.c_int32(0)
        cs = N.Cloud(S.ctypes.data, len(S))
        self._check(self._lib.rst_remove_nans(self._ctx, C.byref(cs), out.ctypes.data, C.byref(n)))
        return out[:n.value].copy()

    def gicp_evaluate(self, src, dst, src_covs, dst_covs, dst_indices, T, huber: float = 0.5, want_residuals: bool = True):
        """GICP residuals / cost / normal equations at pose T (gicp_cost.hpp:40-73, align_gicp.cpp:59-77) on the GPU."""
        S, D = np.ascontiguousarray(src, dtype=np.float32), np.ascontiguousarray(dst, dtype=np.float32)
        cs_, cd_ = np.ascontiguousarray(src_covs, dtype=np.float32), np.ascontiguousarray(dst_covs, dtype=np.float32)
        idx = np.ascontiguousarray(dst_indices, dtype=np.int32)
        pose = pose_to_cm(T)
        res = np.zeros((len(S), 3), dtype=np.float32) if want_residuals else None
        st = N.GicpStats()
        a, b = N.Cloud(S.ctypes.data, len(S)), N.Cloud(D.ctypes.data, len(D))
        self._check(self._lib.rst_gicp_evaluate(self._ctx, C.byref(a), C.byref(b), cs_.ctypes.data, cd_.ctypes.data, idx.ctypes.data,
                                                pose.ctypes.data, huber, res.ctypes.data if res is not None else None, C.byref(st)))
        return res, st

    def gicp_minimize(self, src, dst, src_covs, dst_covs, dst_indices, T0=None, max_iters: int = 32, huber: float = 0.5):
        """The 7-argument ComputeAlignment (align_gicp.cpp:41-117) on the GPU: minimises the robustified GICP cost over the
        pose for given covariances and correspondences, from the seed T0. Returns (pose 4x4, GicpStats at that pose)."""
        S, D = np.ascontiguousarray(src, dtype=np.float32), np.ascontiguousarray(dst, dtype=np.float32)
        cs_, cd_ = np.ascontiguousarray(src_covs, dtype=np.float32), np.ascontiguousarray(dst_covs, dtype=np.float32)
        idx = np.ascontiguousarray(dst_indices, dtype=np.int32)
        pose = pose_to_cm(np.eye(4) if T0 is None else T0).copy()
        st = N.GicpStats()
        a, b = N.Cloud(S.ctypes.data, len(S)), N.Cloud(D.ctypes.data, len(D))
        self._check(self._lib.rst_gicp_minimize(self._ctx, C.byref(a), C.byref(b), cs_.ctypes.data, cd_.ctypes.data, idx.ctypes.data,
                                                max_iters, huber, pose.ctypes.data, C.byref(st)))
        return cm_to_pose(pose), st

    def gicp_align(self, src, dst, max_outer: int = 16, inner_iters: int = 4, huber: float = 0.5, use_gicp_covariances: bool = False,
                   T0=None, grid_cell: float = 0.0):
        """The 3-argument ComputeAlignment (align_gicp.cpp:119-163) on the GPU. Returns (pose 4x4, GicpStats)."""
        S, D = np.ascontiguousarray(src, dtype=np.float32), np.ascontiguousarray(dst, dtype=np.float32)
        pose = pose_to_cm(np.eye(4) if T0 is None else T0).copy()
        st = N.GicpStats()
        a, b = N.Cloud(S.ctypes.data, len(S)), N.Cloud(D.ctypes.data, len(D))
        self._check(self._lib.rst_gicp_align(self._ctx, C.byref(a), C.byref(b), max_outer, inner_iters, huber, 1 if use_gicp_covariances else 0,
                                             grid_cell, pose.ctypes.data, C.byref(st)))
        return cm_to_pose(pose), st

    def icp3d_depth(self, frames: np.ndarray, src_idx, dst_idx, intr, depth_scale: float = 0.001, voxel: float = 0.05,
                    max_iter: int = 128, T0=None, grid_cell: float = 0.1):
        """The reference caller's per-pair sequence from depth frames (rs_replay_app.cpp:229,246-251), on the GPU.
        frames [n,h,w] uint16; returns (ok, poses, mean_cost [n_pairs], cloud sizes [n_frames])."""
        s = np.ascontiguousarray(src_idx, dtype=np.int32)
        d = np.ascontiguousarray(dst_idx, dtype=np.int32)
        n = len(s)
        K = Intrinsics(*intr)
        poses = pose_to_cm(np.broadcast_to(np.eye(4) if T0 is None else T0, (n, 4, 4))).copy()
        res = (N.Icp3dResult * max(n, 1))()
        counts = np.zeros(frames.shape[0], dtype=np.int32)
        self._check(self._lib.rst_icp3d_depth(self._ctx, _frames(frames), frames.shape[0], s.ctypes.data, d.ctypes.data, n,
                                              C.byref(K), depth_scale, voxel, max_iter, grid_cell, poses.ctypes.data,
                                              C.addressof(res), counts.ctypes.data))
        return (np.array([res[i].ok != 0 for i in range(n)]), cm_to_pose(poses),
                np.array([res[i].mean_cost for i in range(n)]), counts)

    def icp3d_read_cloud(self, frame_index: int, n_points: int) -> np.ndarray:
        out = np.empty((n_points, 3), dtype=np.float32)
        self._check(self._lib.rst_icp3d_read_cloud(self._ctx, frame_index, out.ctypes.data, n_points))
        return out

    # ---- staged API ---------------------------------------------------------------------------
    def begin(self, w: int, h: int, intr, params: Params):
        K = Intrinsics(*intr)
        self._check(self._lib.rst_begin(self._ctx, w, h, C.byref(K), C.byref(params)))

    def upload(self, frames: np.ndarray, first_slot: int = 0, rgb: np.ndarray | None = None):
        self._check(self._lib.rst_upload_frames(self._ctx, _frames(frames, rgb), frames.shape[0], first_slot))

    def set_frames_device(self, dev_ptr: int, n: int, row_stride_px: int, frame_stride_px: int, first_slot: int = 0):
        self._check(self._lib.rst_set_frames_device(self._ctx, dev_ptr, n, row_stride_px, frame_stride_px, first_slot))

    def preprocess(self, first_slot: int, n: int):
        self._check(self._lib.rst_preprocess(self._ctx, first_slot, n))

    def align_slots(self, src_slots, dst_slots, T0=None, fetch: bool = True):
        s = np.ascontiguousarray(src_slots, dtype=np.int32)
        d = np.ascontiguousarray(dst_slots, dtype=np.int32)
        n = len(s)
        if not fetch:
            assert T0 is None
            self._check(self._lib.rst_align_slots(self._ctx, s.ctypes.data, d.ctypes.data, n, None, None))
            return None, None
        poses = pose_to_cm(np.broadcast_to(np.eye(4) if T0 is None else T0, (n, 4, 4))).copy()
        stats = (Stats * n)()
        self._check(self._lib.rst_align_slots(self._ctx, s.ctypes.data, d.ctypes.data, n, poses.ctypes.data,
                                              C.addressof(stats)))
        return cm_to_pose(poses), list(stats)

    def sync(self):
        self._check(self._lib.rst_sync(self._ctx))

    def device_results(self):
        dp, ds = C.c_void_p(), C.c_void_p()
        self._check(self._lib.rst_device_results(self._ctx, C.byref(dp), C.byref(ds)))
        return dp.value, ds.value

    def level_info(self, level: int):
        w, h, p = C.c_int32(), C.c_int32(), C.c_int32()
        K = Intrinsics()
        self._check(self._lib.rst_level_info(self._ctx, level, C.byref(w), C.byref(h), C.byref(p), C.byref(K)))
        return w.value, h.value, p.value, (K.fx, K.fy, K.cx, K.cy)

    def read_depth(self, slot: int, level: int) -> np.ndarray:
        w, h, _, _ = self.level_info(level)
        out = np.empty((h, w), dtype=np.uint16)
        self._check(self._lib.rst_read_depth(self._ctx, slot, level, out.ctypes.data))
        return out

    def read_geometry(self, slot: int, level: int) -> np.ndarray:
        w, h, _, _ = self.level_info(level)
        out = np.empty((h, w, 4), dtype=np.float32)
        self._check(self._lib.rst_read_geometry(self._ctx, slot, level, out.ctypes.data))
        return out

    def read_intensity(self, slot: int, level: int) -> np.ndarray:
        w, h, _, _ = self.level_info(level)
        out = np.empty((h, w), dtype=np.float32)
        self._check(self._lib.rst_read_intensity(self._ctx, slot, level, out.ctypes.data))
        return out

    def evaluate(self, src_slot: int, dst_slot: int, level: int, T, want_idx: bool = True):
        w, h, _, _ = self.level_info(level)
        pose = pose_to_cm(T)
        idx = np.empty((h, w), dtype=np.int32) if want_idx else None
        st = Stats()
        self._check(self._lib.rst_evaluate(self._ctx, src_slot, dst_slot, level, pose.ctypes.data,
                                           idx.ctypes.data if want_idx else None, C.byref(st)))
        return idx, st

    def set_stream_split(self, min_pairs: int):
        self._check(self._lib.rst_set_stream_split(self._ctx, min_pairs))

    def set_graph_max_pairs(self, max_pairs: int):
        """Blocking host-frame calls with at most this many pairs replay a CUDA graph (default 8; 0 = never)."""
        self._check(self._lib.rst_set_graph_max_pairs(self._ctx, max_pairs))

    def set_icp3d_cluster(self, ctas_per_pair: int):
        """CTAs per pair of the cloud ICP kernel: 0 = automatic, or 1 / 2 / 4 / 8 / 16."""
        self._check(self._lib.rst_set_icp3d_cluster(self._ctx, ctas_per_pair))

    def set_icp3d_cache(self, gain: float = 1.0, lo_cells: float = 0.05, hi_cells: float = 0.2):
        """Neighbour cache of the cloud ICP kernel (scan margin = clamp(gain * motion, lo, hi) in grid cells); hi_cells = 0: off."""
        self._check(self._lib.rst_set_icp3d_cache(self._ctx, gain, lo_cells, hi_cells))

    def set_icp3d_fixed_point_skip(self, on: bool = True):
        """Cloud ICP: jump over iterations that provably repeat the previous one (default on; results unchanged)."""
        self._check(self._lib.rst_set_icp3d_fixed_point_skip(self._ctx, 1 if on else 0))

    def icp3d_iteration_stats(self):
        """(iterations run, iterations asked for) summed over the pairs of the last cloud-ICP call."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self._lib.rst_icp3d_iteration_stats(self._ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def icp3d_cache_stats(self):
        """(neighbour queries that searched, neighbour queries answered) of the last cloud-ICP call."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self._lib.rst_icp3d_cache_stats(self._ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_schedule(self, schedule: int):
        """0 = fused (one cluster per pair, all iterations in one launch; default), 1 = one launch per iteration."""
        self._check(self._lib.rst_set_schedule(self._ctx, schedule))

    def set_cluster_size(self, tiling: int, ctas_per_pair: int):
        self._check(self._lib.rst_set_cluster_size(self._ctx, tiling, ctas_per_pair))

    def max_active_clusters(self, ctas_per_pair: int) -> int:
        return int(self._lib.rst_max_active_clusters(self._ctx, ctas_per_pair))

    def set_pipeline_chunk(self, frames_per_chunk: int):
        self._check(self._lib.rst_set_pipeline_chunk(self._ctx, frames_per_chunk))

    def copy_results_device(self, d_poses_ptr: int | None, d_stats_ptr: int | None = None):
        self._check(self._lib.rst_copy_results_device(self._ctx, d_poses_ptr, d_stats_ptr))

    def profile_enable(self, on: bool = True):
        self._check(self._lib.rst_profile_enable(self._ctx, 1 if on else 0))

    def profile_read(self) -> "N.Profile":
        p = N.Profile()
        self._check(self._lib.rst_profile_read(self._ctx, C.byref(p)))
        return p

    @property
    def launch_count(self) -> int:
        return int(self._lib.rst_launch_count(self._ctx))


def AlignIcp3d(src: np.ndarray, dst: np.ndarray, max_iter: int = 128, transform=None, aligner: Aligner | None = None):
    """The reference's signature on the GPU: bool AlignIcp3d(src, dst, max_iter, &transform)
    (align_icp.hpp:22-24). Returns (success, transform 4x4)."""
    own = aligner is None
    al = aligner or Aligner(16, 16, 2, 1)
    try:
        ok, T = al.icp3d_pairs([src], [dst], max_iter, T0=transform)
    finally:
        if own:
            al.close()
    return bool(ok[0]), T[0]


def AlignRgbd(src_depth: np.ndarray, dst_depth: np.ndarray, K, params: Params | None = None, transform=None,
              aligner: Aligner | None = None):
    """Frame-based sibling of the reference's AlignIcp3d(src, dst, max_iter, &transform)
    (align_icp.hpp:22-24): returns (success, transform 4x4, stats). `transform` is the initial
    guess (identity if None), src -> dst."""
    h, w = src_depth.shape
    own = aligner is None
    al = aligner or Aligner(w, h, 2, 1)
    try:
        T, st = al.align_pairs(src_depth[None], dst_depth[None], K, params, T0=transform)
    finally:
        if own:
            al.close()
    return st[0].status == N.RST_STATUS_OK, T[0], st[0]
