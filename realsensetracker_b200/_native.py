"""ctypes view of the C ABI in include/rst_align.h (and csrc/rst_synth.h).

There is no CPU fallback: if librst_align.so is missing this module raises at
load time, and every compute call raises RstError when CUDA reports a failure.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIBDIR = Path(__file__).resolve().parent / "_lib"
RST_MAX_LEVELS = 4

RST_OK, RST_ERR_INVALID_ARG, RST_ERR_NO_DEVICE, RST_ERR_CUDA, RST_ERR_CAPACITY, RST_ERR_ALIGNMENT, RST_ERR_ARCH = range(7)
RST_STATUS_OK, RST_STATUS_TOO_FEW, RST_STATUS_DEGENERATE, RST_STATUS_NON_FINITE = 0, 1, 2, 4
RST_ROBUST_NONE, RST_ROBUST_HUBER, RST_ROBUST_GEMAN_MCCLURE = 0, 1, 2
RST_TILING_THROUGHPUT, RST_TILING_LATENCY = 0, 1

ERR_NAMES = {0: "RST_OK", 1: "RST_ERR_INVALID_ARG", 2: "RST_ERR_NO_DEVICE", 3: "RST_ERR_CUDA",
             4: "RST_ERR_CAPACITY", 5: "RST_ERR_ALIGNMENT", 6: "RST_ERR_ARCH"}


class RstError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Intrinsics(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class Frame(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("rgb", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32),
                ("depth_stride_bytes", C.c_int32), ("rgb_stride_bytes", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("num_levels", C.c_int32), ("iters", C.c_int32 * RST_MAX_LEVELS),
                ("depth_scale", C.c_float), ("z_min", C.c_float), ("z_max", C.c_float),
                ("dist_max", C.c_float), ("normal_cos_min", C.c_float), ("normal_depth_tol", C.c_float),
                ("pyr_depth_tol", C.c_int32), ("robust_kind", C.c_int32), ("robust_scale", C.c_float),
                ("min_count", C.c_int32), ("damping", C.c_float), ("photo_weight", C.c_float),
                ("tiling", C.c_int32), ("converge_eps", C.c_float), ("reserved", C.c_int32 * 2)]


class Stats(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations", C.c_int32), ("count", C.c_int32), ("rmse", C.c_float),
                ("any_status", C.c_int32), ("failed_iterations", C.c_int32), ("sum_wr2", C.c_double), ("A", C.c_double * 21), ("b", C.c_double * 6)]


class GicpStats(C.Structure):
    _fields_ = [("cost", C.c_double), ("A", C.c_double * 21), ("b", C.c_double * 6), ("count", C.c_int32), ("reserved", C.c_int32)]


class Cloud(C.Structure):
    _fields_ = [("xyz", C.c_void_p), ("n", C.c_int32)]


class Icp3dResult(C.Structure):
    _fields_ = [("ok", C.c_int32), ("iterations", C.c_int32), ("mean_cost", C.c_float), ("mu", C.c_float),
                ("cov", C.c_double * 9)]


class Profile(C.Structure):
    _fields_ = [("ms_preprocess", C.c_float * RST_MAX_LEVELS), ("ms_icp", C.c_float * RST_MAX_LEVELS),
                ("launches_preprocess", C.c_int32 * RST_MAX_LEVELS), ("launches_icp", C.c_int32 * RST_MAX_LEVELS),
                ("frames_preprocessed", C.c_int64 * RST_MAX_LEVELS), ("pairs_iterated", C.c_int64 * RST_MAX_LEVELS),
                ("ms_icp_fused", C.c_float), ("launches_icp_fused", C.c_int32), ("pair_iterations_fused", C.c_int64)]


# every symbol include/rst_align.h declares (tests check the library exports all of them)
ALIGN_SYMBOLS = [
    "rst_ctx_create", "rst_ctx_destroy", "rst_last_error", "rst_last_create_error", "rst_abi_version",
    "rst_params_default", "rst_align_pairs", "rst_align_sequence", "rst_begin", "rst_upload_frames",
    "rst_set_frames_device", "rst_preprocess", "rst_align_slots", "rst_device_results", "rst_sync",
    "rst_level_info", "rst_read_depth", "rst_read_geometry", "rst_read_intensity", "rst_evaluate", "rst_launch_count",
    "rst_copy_results_device", "rst_profile_enable", "rst_profile_read", "rst_set_pipeline_chunk", "rst_set_stream_split", "rst_align_pairs_async", "rst_align_sequence_async", "rst_wait", "rst_icp3d_pairs", "rst_solve_kabsch", "rst_cloud_normals", "rst_icp3d_depth", "rst_icp3d_read_cloud",
    "rst_set_schedule", "rst_set_cluster_size", "rst_max_active_clusters",
    "rst_find_correspondences", "rst_cloud_covariances", "rst_downsample_voxel", "rst_remove_nans", "rst_cloud_centroid", "rst_orient_normals", "rst_cloud_extents", "rst_tree_create", "rst_tree_query", "rst_tree_size", "rst_tree_destroy",
    "rst_gicp_evaluate", "rst_gicp_minimize", "rst_gicp_align", "rst_set_graph_max_pairs", "rst_set_icp3d_cluster", "rst_set_icp3d_cache", "rst_icp3d_cache_stats", "rst_set_icp3d_fixed_point_skip", "rst_icp3d_iteration_stats",
]

_align = None
_synth = None


def _load(name: str) -> C.CDLL:
    import os
    path = LIBDIR / name
    if name == "librst_align.so" and os.environ.get("RST_ALIGN_LIB"):  # kernel-variant experiments
        path = Path(os.environ["RST_ALIGN_LIB"])
    if not path.exists():
        raise ImportError(
            f"{path} is missing — build it with `python -m realsensetracker_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for the alignment path.")
    return C.CDLL(str(path))


def align_lib() -> C.CDLL:
    global _align
    if _align is None:
        lib = _load("librst_align.so")
        P = C.POINTER
        lib.rst_ctx_create.argtypes = [C.c_int32] * 5 + [C.c_void_p, P(C.c_void_p)]
        lib.rst_ctx_create.restype = C.c_int32
        lib.rst_ctx_destroy.argtypes = [C.c_void_p]
        lib.rst_ctx_destroy.restype = None
        lib.rst_last_error.argtypes = [C.c_void_p]
        lib.rst_last_error.restype = C.c_char_p
        lib.rst_last_create_error.argtypes = []
        lib.rst_last_create_error.restype = C.c_char_p
        lib.rst_abi_version.restype = C.c_int32
        lib.rst_params_default.argtypes = [P(Params)]
        lib.rst_params_default.restype = None
        lib.rst_align_pairs.argtypes = [C.c_void_p, P(Frame), P(Frame), C.c_int32, P(Intrinsics), P(Params),
                                        C.c_void_p, C.c_void_p]
        lib.rst_align_pairs.restype = C.c_int32
        lib.rst_align_sequence.argtypes = [C.c_void_p, P(Frame), C.c_int32, P(Intrinsics), P(Params),
                                           C.c_void_p, C.c_void_p]
        lib.rst_align_sequence.restype = C.c_int32
        lib.rst_align_pairs_async.argtypes = [C.c_void_p, P(Frame), P(Frame), C.c_int32, P(Intrinsics), P(Params), C.c_void_p]
        lib.rst_align_pairs_async.restype = C.c_int32
        lib.rst_align_sequence_async.argtypes = [C.c_void_p, P(Frame), C.c_int32, P(Intrinsics), P(Params), C.c_void_p]
        lib.rst_align_sequence_async.restype = C.c_int32
        lib.rst_wait.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rst_wait.restype = C.c_int32
        lib.rst_icp3d_pairs.argtypes = [C.c_void_p, P(Cloud), P(Cloud), C.c_int32, C.c_int32, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rst_icp3d_pairs.restype = C.c_int32
        lib.rst_solve_kabsch.argtypes = [C.c_void_p, P(Cloud), P(Cloud), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, P(C.c_int32)]
        lib.rst_solve_kabsch.restype = C.c_int32
        lib.rst_cloud_normals.argtypes = [C.c_void_p, P(Cloud), C.c_int32, C.c_void_p, C.c_float, C.c_void_p]
        lib.rst_cloud_normals.restype = C.c_int32
        lib.rst_icp3d_depth.argtypes = [C.c_void_p, P(Frame), C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, P(Intrinsics),
                                        C.c_float, C.c_float, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rst_icp3d_depth.restype = C.c_int32
        lib.rst_find_correspondences.argtypes = [C.c_void_p, P(Cloud), P(Cloud), C.c_float, C.c_void_p, C.c_void_p]
        lib.rst_find_correspondences.restype = C.c_int32
        lib.rst_cloud_covariances.argtypes = [C.c_void_p, P(Cloud), C.c_int32, C.c_float, C.c_void_p]
        lib.rst_cloud_covariances.restype = C.c_int32
        lib.rst_downsample_voxel.argtypes = [C.c_void_p, P(Cloud), C.c_float, C.c_void_p, P(C.c_int32)]
        lib.rst_downsample_voxel.restype = C.c_int32
        lib.rst_remove_nans.argtypes = [C.c_void_p, P(Cloud), C.c_void_p, P(C.c_int32)]
        lib.rst_remove_nans.restype = C.c_int32
        lib.rst_cloud_centroid.argtypes = [C.c_void_p, P(Cloud), C.c_void_p]
        lib.rst_cloud_centroid.restype = C.c_int32
        lib.rst_tree_create.argtypes = [C.c_void_p, P(Cloud), C.c_float, P(C.c_void_p)]
        lib.rst_tree_create.restype = C.c_int32
        lib.rst_tree_query.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        lib.rst_tree_query.restype = C.c_int32
        lib.rst_tree_size.argtypes = [C.c_void_p]
        lib.rst_tree_size.restype = C.c_int32
        lib.rst_tree_destroy.argtypes = [C.c_void_p]
        lib.rst_tree_destroy.restype = None
        lib.rst_cloud_extents.argtypes = [C.c_void_p, P(Cloud), C.c_void_p, C.c_void_p]
        lib.rst_cloud_extents.restype = C.c_int32
        lib.rst_orient_normals.argtypes = [C.c_void_p, P(Cloud), C.c_void_p, C.c_void_p]
        lib.rst_orient_normals.restype = C.c_int32
        lib.rst_gicp_minimize.argtypes = [C.c_void_p, P(Cloud), P(Cloud), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float,
                                          C.c_void_p, P(GicpStats)]
        lib.rst_gicp_minimize.restype = C.c_int32
        lib.rst_gicp_evaluate.argtypes = [C.c_void_p, P(Cloud), P(Cloud), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                          C.c_void_p, P(GicpStats)]
        lib.rst_gicp_evaluate.restype = C.c_int32
        lib.rst_gicp_align.argtypes = [C.c_void_p, P(Cloud), P(Cloud), C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_float,
                                       C.c_void_p, P(GicpStats)]
        lib.rst_gicp_align.restype = C.c_int32
        lib.rst_icp3d_read_cloud.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        lib.rst_icp3d_read_cloud.restype = C.c_int32
        lib.rst_begin.argtypes = [C.c_void_p, C.c_int32, C.c_int32, P(Intrinsics), P(Params)]
        lib.rst_begin.restype = C.c_int32
        lib.rst_upload_frames.argtypes = [C.c_void_p, P(Frame), C.c_int32, C.c_int32]
        lib.rst_upload_frames.restype = C.c_int32
        lib.rst_set_frames_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32]
        lib.rst_set_frames_device.restype = C.c_int32
        lib.rst_preprocess.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        lib.rst_preprocess.restype = C.c_int32
        lib.rst_align_slots.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.rst_align_slots.restype = C.c_int32
        lib.rst_device_results.argtypes = [C.c_void_p, P(C.c_void_p), P(C.c_void_p)]
        lib.rst_device_results.restype = C.c_int32
        lib.rst_sync.argtypes = [C.c_void_p]
        lib.rst_sync.restype = C.c_int32
        lib.rst_level_info.argtypes = [C.c_void_p, C.c_int32, P(C.c_int32), P(C.c_int32), P(C.c_int32), P(Intrinsics)]
        lib.rst_level_info.restype = C.c_int32
        lib.rst_read_depth.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        lib.rst_read_depth.restype = C.c_int32
        lib.rst_read_geometry.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        lib.rst_read_geometry.restype = C.c_int32
        lib.rst_read_intensity.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        lib.rst_read_intensity.restype = C.c_int32
        lib.rst_evaluate.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, P(Stats)]
        lib.rst_evaluate.restype = C.c_int32
        lib.rst_launch_count.argtypes = [C.c_void_p]
        lib.rst_launch_count.restype = C.c_int64
        lib.rst_copy_results_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rst_copy_results_device.restype = C.c_int32
        lib.rst_set_stream_split.argtypes = [C.c_void_p, C.c_int32]
        lib.rst_set_stream_split.restype = C.c_int32
        lib.rst_set_icp3d_cluster.argtypes = [C.c_void_p, C.c_int32]
        lib.rst_set_icp3d_cluster.restype = C.c_int32
        lib.rst_set_icp3d_cache.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        lib.rst_set_icp3d_cache.restype = C.c_int32
        lib.rst_icp3d_cache_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.rst_icp3d_cache_stats.restype = C.c_int32
        lib.rst_set_icp3d_fixed_point_skip.argtypes = [C.c_void_p, C.c_int32]
        lib.rst_set_icp3d_fixed_point_skip.restype = C.c_int32
        lib.rst_icp3d_iteration_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.rst_icp3d_iteration_stats.restype = C.c_int32
        lib.rst_set_graph_max_pairs.argtypes = [C.c_void_p, C.c_int32]
        lib.rst_set_graph_max_pairs.restype = C.c_int32
        lib.rst_set_schedule.argtypes = [C.c_void_p, C.c_int32]
        lib.rst_set_schedule.restype = C.c_int32
        lib.rst_set_cluster_size.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        lib.rst_set_cluster_size.restype = C.c_int32
        lib.rst_max_active_clusters.argtypes = [C.c_void_p, C.c_int32]
        lib.rst_max_active_clusters.restype = C.c_int32
        lib.rst_set_pipeline_chunk.argtypes = [C.c_void_p, C.c_int32]
        lib.rst_set_pipeline_chunk.restype = C.c_int32
        lib.rst_profile_enable.argtypes = [C.c_void_p, C.c_int32]
        lib.rst_profile_enable.restype = C.c_int32
        lib.rst_profile_read.argtypes = [C.c_void_p, P(Profile)]
        lib.rst_profile_read.restype = C.c_int32
        _align = lib
    return _align


class SynthScene(C.Structure):
    _fields_ = [("room_lo", C.c_double * 3), ("room_hi", C.c_double * 3), ("n_spheres", C.c_int32),
                ("sphere_c", (C.c_double * 3) * 8), ("sphere_r", C.c_double * 8), ("n_boxes", C.c_int32),
                ("box_c", (C.c_double * 3) * 4), ("box_h", (C.c_double * 3) * 4), ("box_yaw", C.c_double * 4)]


class SynthNoise(C.Structure):
    _fields_ = [("sigma_lsb_at_1m", C.c_double), ("p_invalid_pixel", C.c_double), ("p_invalid_block", C.c_double)]


def synth_lib() -> C.CDLL:
    global _synth
    if _synth is None:
        lib = _load("librst_synth.so")
        lib.rst_synth_scene_default.argtypes = [C.c_uint64, C.POINTER(SynthScene)]
        lib.rst_synth_scene_default.restype = None
        lib.rst_synth_render.argtypes = [C.POINTER(SynthScene), C.c_void_p] + [C.c_double] * 4 + \
            [C.c_int32, C.c_int32, C.c_double, C.POINTER(SynthNoise), C.c_uint64, C.c_void_p, C.c_int32, C.c_void_p]
        lib.rst_synth_render.restype = C.c_int64
        _synth = lib
    return _synth
