"""ctypes bindings of oracle/_build/liboracle.so (Oracle-N + Oracle-R). Checker only."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liboracle.so"
RST_MAX_LEVELS = 4


class Intrinsics(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


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


class Level(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("ifx", C.c_float), ("ify", C.c_float)]


def build(force: bool = False) -> Path:
    """Compiles the oracles with oracle/Makefile (gcc; no GPU involved)."""
    if force:
        subprocess.run(["make", "-C", str(HERE), "clean"], check=True, stdout=subprocess.DEVNULL)
    res = subprocess.run(["make", "-C", str(HERE)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout)
    return LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        _lib = C.CDLL(str(LIB))
        _lib.or_align_icp3d.restype = C.c_int32
        _lib.or_align_depth_pair.restype = C.c_int32
        _lib.or_solve_kabsch.restype = C.c_int32
        _lib.or_remove_nans.restype = C.c_int32
        _lib.or_downsample_voxel.restype = C.c_int32
        _lib.on_solve.restype = C.c_int32
        _lib.on_align_pair.restype = C.c_int32
        _lib.on_align_pair_rgbd.restype = C.c_int32
    return _lib


def default_params(**kw) -> Params:
    """Same defaults as rst_params_default() (include/rst_align.h); kept in sync by a test."""
    p = Params()
    p.num_levels = 3
    p.iters[0], p.iters[1], p.iters[2], p.iters[3] = 10, 5, 4, 0
    p.depth_scale = 0.001
    p.z_min, p.z_max = 0.1, 10.0
    p.dist_max = 0.2
    p.normal_cos_min = -2.0
    p.normal_depth_tol = 0.05
    p.pyr_depth_tol = 100
    p.robust_kind = 0
    p.robust_scale = 0.02
    p.min_count = 16
    p.damping = 0.0
    p.photo_weight = 0.0
    p.tiling = 0
    p.converge_eps = 0.0
    for k, v in kw.items():
        if k == "iters":
            for i, x in enumerate(v):
                p.iters[i] = x
        else:
            setattr(p, k, v)
    return p


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def pose_to_cm(T) -> np.ndarray:
    """4x4 (numpy row-major view) -> 16 floats column-major."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(16)


def cm_to_pose(p) -> np.ndarray:
    return np.asarray(p, dtype=np.float64).reshape(4, 4).T.copy()


# ------------------------------------------------------------------ Oracle-N
def level_info(intr, w, h, level) -> Level:
    K = Intrinsics(*intr)
    L = Level()
    lib().on_level_info(C.byref(K), C.c_int32(w), C.c_int32(h), C.c_int32(level), C.byref(L))
    return L


def pyr_down(depth: np.ndarray, tol: int) -> np.ndarray:
    h, w = depth.shape
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    out = np.empty((h // 2, w // 2), dtype=np.uint16)
    lib().on_pyr_down(d.ctypes.data_as(C.c_void_p), C.c_int32(w), C.c_int32(h), C.c_int32(tol),
                      out.ctypes.data_as(C.c_void_p))
    return out


def geometry(depth: np.ndarray, L: Level, P: Params) -> np.ndarray:
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    G = np.empty((L.h, L.w, 4), dtype=np.float32)
    lib().on_geometry(d.ctypes.data_as(C.c_void_p), C.byref(L), C.byref(P), G.ctypes.data_as(C.c_void_p))
    return G


def evaluate(src_depth, src_G, dst_G, L: Level, P: Params, T, want_idx=True):
    d = np.ascontiguousarray(src_depth, dtype=np.uint16)
    pose = pose_to_cm(T)
    idx = np.empty((L.h, L.w), dtype=np.int32) if want_idx else None
    st = Stats()
    lib().on_evaluate(d.ctypes.data_as(C.c_void_p),
                      src_G.ctypes.data_as(C.c_void_p) if src_G is not None else None,
                      dst_G.ctypes.data_as(C.c_void_p), C.byref(L), C.byref(P),
                      pose.ctypes.data_as(C.c_void_p),
                      idx.ctypes.data_as(C.c_void_p) if want_idx else None, C.byref(st))
    return idx, st


def align_pair(src: np.ndarray, dst: np.ndarray, intr, P: Params, T0=None):
    """Full Oracle-N alignment. Returns (T 4x4 float64 view of the fp32 result, Stats)."""
    h, w = src.shape
    s = np.ascontiguousarray(src, dtype=np.uint16)
    d = np.ascontiguousarray(dst, dtype=np.uint16)
    K = Intrinsics(*intr)
    pose = pose_to_cm(np.eye(4) if T0 is None else T0)
    st = Stats()
    lib().on_align_pair(s.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.c_int32(w), C.c_int32(h),
                        C.byref(K), C.byref(P), pose.ctypes.data_as(C.c_void_p), C.byref(st))
    return cm_to_pose(pose), st


def intensity(rgb: np.ndarray) -> np.ndarray:
    h, w, _ = rgb.shape
    c = np.ascontiguousarray(rgb, dtype=np.uint8)
    out = np.empty((h, w), dtype=np.float32)
    lib().on_intensity(c.ctypes.data_as(C.c_void_p), C.c_int32(w), C.c_int32(h), out.ctypes.data_as(C.c_void_p))
    return out


def intensity_down(I: np.ndarray) -> np.ndarray:
    h, w = I.shape
    a = np.ascontiguousarray(I, dtype=np.float32)
    out = np.empty((h // 2, w // 2), dtype=np.float32)
    lib().on_intensity_down(a.ctypes.data_as(C.c_void_p), C.c_int32(w), C.c_int32(h), out.ctypes.data_as(C.c_void_p))
    return out


def evaluate_photo(src_depth, src_G, dst_G, src_I, dst_I, L: Level, P: Params, T, want_idx=True):
    d = np.ascontiguousarray(src_depth, dtype=np.uint16)
    si, di = np.ascontiguousarray(src_I, dtype=np.float32), np.ascontiguousarray(dst_I, dtype=np.float32)
    pose = pose_to_cm(T)
    idx = np.empty((L.h, L.w), dtype=np.int32) if want_idx else None
    st = Stats()
    lib().on_evaluate_photo(d.ctypes.data_as(C.c_void_p),
                            src_G.ctypes.data_as(C.c_void_p) if src_G is not None else None,
                            dst_G.ctypes.data_as(C.c_void_p), si.ctypes.data_as(C.c_void_p), di.ctypes.data_as(C.c_void_p),
                            C.byref(L), C.byref(P), pose.ctypes.data_as(C.c_void_p),
                            idx.ctypes.data_as(C.c_void_p) if want_idx else None, C.byref(st))
    return idx, st


def align_pair_rgbd(src, dst, src_rgb, dst_rgb, intr, P: Params, T0=None):
    h, w = src.shape
    s, d = np.ascontiguousarray(src, dtype=np.uint16), np.ascontiguousarray(dst, dtype=np.uint16)
    sc, dc = np.ascontiguousarray(src_rgb, dtype=np.uint8), np.ascontiguousarray(dst_rgb, dtype=np.uint8)
    K = Intrinsics(*intr)
    pose = pose_to_cm(np.eye(4) if T0 is None else T0)
    st = Stats()
    lib().on_align_pair_rgbd(s.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                             dc.ctypes.data_as(C.c_void_p), C.c_int32(w), C.c_int32(h), C.byref(K), C.byref(P),
                             pose.ctypes.data_as(C.c_void_p), C.byref(st))
    return cm_to_pose(pose), st


def solve(A21, b6, count, P: Params):
    A = np.ascontiguousarray(A21, dtype=np.float64)
    b = np.ascontiguousarray(b6, dtype=np.float64)
    xi = np.zeros(6)
    rc = lib().on_solve(A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_int32(count), C.byref(P),
                        xi.ctypes.data_as(C.c_void_p))
    return rc, xi


def pose_update(xi, T):
    Rt = np.concatenate([np.asarray(T)[:3, :3].reshape(9), np.asarray(T)[:3, 3]]).astype(np.float64)
    x = np.ascontiguousarray(xi, dtype=np.float64)
    lib().on_pose_update(x.ctypes.data_as(C.c_void_p), Rt.ctypes.data_as(C.c_void_p))
    out = np.eye(4)
    out[:3, :3] = Rt[:9].reshape(3, 3)
    out[:3, 3] = Rt[9:]
    return out


# ------------------------------------------------------------------ Oracle-R
def centroid(pts):
    p = _f32(pts)
    c = np.zeros(3, dtype=np.float32)
    lib().or_centroid(p.ctypes.data_as(C.c_void_p), C.c_int32(len(p)), c.ctypes.data_as(C.c_void_p))
    return c


def extents(pts):
    """ComputeExtents (point_cloud_utils.cpp:26-32): (lo [3], hi [3]); Eigen's empty box for an empty cloud."""
    p = _f32(pts).reshape(-1, 3)
    fmax = np.finfo(np.float32).max
    if len(p) == 0:
        return np.full(3, fmax, np.float32), np.full(3, -fmax, np.float32)
    return p.min(axis=0), p.max(axis=0)


def orient_normals(pts, viewpoint, normals):
    """OrientNormals (point_cloud_utils.cpp:205-216): normal i is negated where (p_i - viewpoint) . n_i > 0; fp32, the
    three products summed left to right without contraction (how the oracle's and the compiled reference's build
    evaluate Eigen's dot). Returns a copy."""
    p, n = _f32(pts).reshape(-1, 3), np.array(normals, dtype=np.float32, order="C", copy=True).reshape(-1, 3)
    ray = p - _f32(viewpoint).reshape(1, 3)
    dot = (ray[:, 0] * n[:, 0] + ray[:, 1] * n[:, 1]) + ray[:, 2] * n[:, 2]
    n[dot > 0] *= np.float32(-1)
    return n


def remove_nans(pts):
    p = _f32(pts)
    out = np.empty_like(p)
    m = lib().or_remove_nans(p.ctypes.data_as(C.c_void_p), C.c_int32(len(p)), out.ctypes.data_as(C.c_void_p))
    return out[:m].copy()


def downsample_voxel(pts, voxel):
    p = _f32(pts)
    out = np.empty_like(p)
    m = lib().or_downsample_voxel(p.ctypes.data_as(C.c_void_p), C.c_int32(len(p)), C.c_float(voxel),
                                  out.ctypes.data_as(C.c_void_p))
    return out[:m].copy()


def nn(dst, queries, leaf=16):
    d, q = _f32(dst), _f32(queries)
    idx = np.empty(len(q), dtype=np.int32)
    d2 = np.empty(len(q), dtype=np.float32)
    lib().or_nn(d.ctypes.data_as(C.c_void_p), C.c_int32(len(d)), q.ctypes.data_as(C.c_void_p), C.c_int32(len(q)),
                C.c_int32(leaf), idx.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p))
    return idx, d2


def solve_kabsch(src, dst, pairs, weights=None):
    s, d = _f32(src), _f32(dst)
    pr = np.ascontiguousarray(pairs, dtype=np.int32)
    w = _f32(weights) if weights is not None else None
    T = np.zeros(16, dtype=np.float32)
    ok = lib().or_solve_kabsch(s.ctypes.data_as(C.c_void_p), C.c_int32(len(s)), d.ctypes.data_as(C.c_void_p),
                               C.c_int32(len(d)), pr.ctypes.data_as(C.c_void_p), C.c_int32(len(pr)),
                               w.ctypes.data_as(C.c_void_p) if w is not None else None,
                               T.ctypes.data_as(C.c_void_p))
    return bool(ok), cm_to_pose(T)


def align_icp3d(src, dst, max_iter=128, T0=None, details=False):
    """Oracle-R AlignIcp3d (align_icp.cpp:73-167). Returns (ok, T[, extras])."""
    s, d = _f32(src), _f32(dst)
    T = pose_to_cm(np.eye(4) if T0 is None else T0)
    mc = C.c_float(0)
    nbrs = np.empty(len(s), dtype=np.int32)
    wts = np.empty(len(s), dtype=np.float32)
    cov = np.zeros(9)
    ok = lib().or_align_icp3d(s.ctypes.data_as(C.c_void_p), C.c_int32(len(s)), d.ctypes.data_as(C.c_void_p),
                              C.c_int32(len(d)), C.c_int32(max_iter), T.ctypes.data_as(C.c_void_p), C.byref(mc),
                              nbrs.ctypes.data_as(C.c_void_p), wts.ctypes.data_as(C.c_void_p),
                              cov.ctypes.data_as(C.c_void_p))
    if details:
        return bool(ok), cm_to_pose(T), dict(mean_cost=mc.value, nbrs=nbrs, weights=wts, cov=cov.reshape(3, 3))
    return bool(ok), cm_to_pose(T)


def backproject(depth, intr, depth_scale=0.001):
    h, w = depth.shape
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    out = np.empty((h * w, 3), dtype=np.float32)
    fx, fy, cx, cy = intr
    lib().or_backproject(d.ctypes.data_as(C.c_void_p), C.c_int32(w), C.c_int32(h), C.c_float(fx), C.c_float(fy),
                         C.c_float(cx), C.c_float(cy), C.c_float(depth_scale), out.ctypes.data_as(C.c_void_p))
    return out


def align_depth_pair(src, dst, intr, depth_scale=0.001, voxel=0.05, max_iter=128, T0=None):
    """The reference caller's per-pair sequence (rs_replay_app.cpp:229,246-251) on depth frames."""
    h, w = src.shape
    s = np.ascontiguousarray(src, dtype=np.uint16)
    d = np.ascontiguousarray(dst, dtype=np.uint16)
    fx, fy, cx, cy = intr
    T = pose_to_cm(np.eye(4) if T0 is None else T0)
    mc = C.c_float(0)
    ns, nd = C.c_int32(0), C.c_int32(0)
    ok = lib().or_align_depth_pair(s.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.c_int32(w),
                                   C.c_int32(h), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                                   C.c_float(depth_scale), C.c_float(voxel), C.c_int32(max_iter),
                                   T.ctypes.data_as(C.c_void_p), C.byref(mc), C.byref(ns), C.byref(nd))
    return bool(ok), cm_to_pose(T), dict(mean_cost=mc.value, n_src=ns.value, n_dst=nd.value)


def align_depth_pairs(src, dst, intr, depth_scale=0.001, voxel=0.05, max_iter=128, n_threads=1):
    """Batch of pairs, one pair per OpenMP thread. src/dst: [n,h,w] uint16."""
    n, h, w = src.shape
    s = np.ascontiguousarray(src, dtype=np.uint16)
    d = np.ascontiguousarray(dst, dtype=np.uint16)
    fx, fy, cx, cy = intr
    T = np.tile(pose_to_cm(np.eye(4)), (n, 1))
    ok = np.zeros(n, dtype=np.int32)
    lib().or_align_depth_pairs(s.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.c_int32(n),
                               C.c_int32(w), C.c_int32(h), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                               C.c_float(cy), C.c_float(depth_scale), C.c_float(voxel), C.c_int32(max_iter),
                               C.c_int32(n_threads), T.ctypes.data_as(C.c_void_p), ok.ctypes.data_as(C.c_void_p))
    return ok.astype(bool), np.stack([cm_to_pose(t) for t in T])


# ------------------------------------------------------------------ compiled reference (oracle/_ref)
REF_LIB = HERE / "_ref" / "libref.so"
REFERENCE_TREE = Path("/root/reference/rs_tracker")
_ref = None


def build_ref() -> Path | None:
    """Compiles the reference's own align_icp.cpp + point_cloud_utils.cpp (unmodified, from
    /root/reference) against oracle/shim into oracle/_ref/libref.so. Returns None when the reference
    tree is not present (GPU box) and no prebuilt library travelled with the snapshot."""
    if REFERENCE_TREE.exists():
        res = subprocess.run(["make", "-C", str(HERE), "ref"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError("reference build failed:\n" + res.stdout)
    return REF_LIB if REF_LIB.exists() else None


def ref_lib():
    global _ref
    if _ref is None:
        p = build_ref()
        if p is None:
            return None
        _ref = C.CDLL(str(p))
    return _ref


def ref_align_icp3d(src, dst, max_iter=128, T0=None):
    s, d = _f32(src), _f32(dst)
    T = pose_to_cm(np.eye(4) if T0 is None else T0)
    ok = ref_lib().ref_align_icp3d(s.ctypes.data_as(C.c_void_p), C.c_int32(len(s)), d.ctypes.data_as(C.c_void_p),
                                   C.c_int32(len(d)), C.c_int32(max_iter), T.ctypes.data_as(C.c_void_p))
    return bool(ok), cm_to_pose(T)


def ref_solve_kabsch(src, dst, pairs, weights=None):
    s, d = _f32(src), _f32(dst)
    pr = np.ascontiguousarray(pairs, dtype=np.int32)
    w = _f32(weights) if weights is not None else None
    T = np.zeros(16, dtype=np.float32)
    ok = ref_lib().ref_solve_kabsch(s.ctypes.data_as(C.c_void_p), C.c_int32(len(s)), d.ctypes.data_as(C.c_void_p),
                                    C.c_int32(len(d)), pr.ctypes.data_as(C.c_void_p), C.c_int32(len(pr)),
                                    w.ctypes.data_as(C.c_void_p) if w is not None else None, T.ctypes.data_as(C.c_void_p))
    return bool(ok), cm_to_pose(T)


def ref_centroid(pts):
    p = _f32(pts)
    c = np.zeros(3, dtype=np.float32)
    ref_lib().ref_centroid(p.ctypes.data_as(C.c_void_p), C.c_int32(len(p)), c.ctypes.data_as(C.c_void_p))
    return c


def ref_remove_nans(pts):
    p = _f32(pts)
    out = np.empty_like(p)
    m = ref_lib().ref_remove_nans(p.ctypes.data_as(C.c_void_p), C.c_int32(len(p)), out.ctypes.data_as(C.c_void_p))
    return out[:m].copy()


def ref_downsample_voxel(pts, voxel):
    p = _f32(pts)
    out = np.empty_like(p)
    m = ref_lib().ref_downsample_voxel(p.ctypes.data_as(C.c_void_p), C.c_int32(len(p)), C.c_float(voxel),
                                       out.ctypes.data_as(C.c_void_p))
    return out[:m].copy()


def ref_find_correspondences(dst, src):
    d, s = _f32(dst), _f32(src)
    idx = np.empty(len(s), dtype=np.int32)
    d2 = np.empty(len(s), dtype=np.float32)
    ref_lib().ref_find_correspondences(d.ctypes.data_as(C.c_void_p), C.c_int32(len(d)), s.ctypes.data_as(C.c_void_p),
                                       C.c_int32(len(s)), idx.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p))
    return idx, d2


def ref_covariances(pts, use_gicp=False):
    """ComputeCovariances (point_cloud_utils.cpp:100-161) through the compiled reference: [n,3,3] float32."""
    p = _f32(pts)
    out = np.empty((len(p), 3, 3), dtype=np.float32)
    ref_lib().ref_covariances(p.ctypes.data_as(C.c_void_p), C.c_int32(len(p)), C.c_int32(1 if use_gicp else 0),
                              out.ctypes.data_as(C.c_void_p))
    return out


def ref_normals(pts, k=16, viewpoint=(0.0, 0.0, 0.0)):
    p = _f32(pts)
    vp = _f32(viewpoint)
    out = np.empty_like(p)
    ref_lib().ref_normals(p.ctypes.data_as(C.c_void_p), C.c_int32(len(p)), C.c_int32(k), vp.ctypes.data_as(C.c_void_p),
                          out.ctypes.data_as(C.c_void_p))
    return out


def ref_extents(pts):
    """ComputeExtents (point_cloud_utils.cpp:26-32) through the compiled reference: (lo [3], hi [3]) float32."""
    p = _f32(pts)
    lo, hi = np.empty(3, dtype=np.float32), np.empty(3, dtype=np.float32)
    ref_lib().ref_extents(p.ctypes.data_as(C.c_void_p), C.c_int32(len(p)), lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p))
    return lo, hi


def ref_orient_normals(pts, viewpoint, normals):
    """OrientNormals (point_cloud_utils.cpp:205-216) through the compiled reference: a flipped copy of `normals`."""
    p = _f32(pts)
    vp = _f32(viewpoint)
    out = np.array(normals, dtype=np.float32, order="C", copy=True)
    ref_lib().ref_orient_normals(p.ctypes.data_as(C.c_void_p), C.c_int32(len(p)), vp.ctypes.data_as(C.c_void_p),
                                 out.ctypes.data_as(C.c_void_p))
    return out


def ref_align_depth_pairs(src, dst, intr, depth_scale=0.001, voxel=0.05, max_iter=128, n_threads=1):
    """Batch of depth pairs through the reference's own RemoveNans / DownsampleVoxel / AlignIcp3d."""
    n, h, w = src.shape
    s = np.ascontiguousarray(src, dtype=np.uint16)
    d = np.ascontiguousarray(dst, dtype=np.uint16)
    fx, fy, cx, cy = intr
    T = np.tile(pose_to_cm(np.eye(4)), (n, 1))
    ok = np.zeros(n, dtype=np.int32)
    ref_lib().ref_align_depth_pairs(s.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.c_int32(n), C.c_int32(w),
                                    C.c_int32(h), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                                    C.c_float(depth_scale), C.c_float(voxel), C.c_int32(max_iter), C.c_int32(n_threads),
                                    T.ctypes.data_as(C.c_void_p), ok.ctypes.data_as(C.c_void_p))
    return ok.astype(bool), np.stack([cm_to_pose(t) for t in T])


# ---------------------------------------------------------------------------------------------
# GICP plane-to-plane residual: float64 numpy restatement (TEST INFRASTRUCTURE, like everything in oracle/).
#   gicp_cost.hpp:48-70   delta = R*src + t - dst;  cov = dst_cov + R*src_cov*R^T;  rsqrt_cov = V diag(eig^-1/2) V^T;
#                         residual = rsqrt_cov * delta
#   align_gicp.cpp:59-77  one residual block per correspondence (src i -> dst dst_indices[i]), ceres::HuberLoss(0.5)
#   align_gicp.cpp:113    returns summary.final_cost = 1/2 * sum rho(|residual|^2)
# The reference differentiates through rsqrt_cov with ceres autodiff; the Gauss-Newton normal equations returned here
# hold the combined covariance at the current rotation (J = rsqrt_cov [ -[p']x | I ], left perturbation, omega first).
# ---------------------------------------------------------------------------------------------
def gicp_evaluate(src, dst, src_covs, dst_covs, dst_indices, T, huber=0.5):
    s = np.asarray(src, dtype=np.float64); d = np.asarray(dst, dtype=np.float64)
    Cs = np.asarray(src_covs, dtype=np.float64).reshape(-1, 3, 3); Cd = np.asarray(dst_covs, dtype=np.float64).reshape(-1, 3, 3)
    idx = np.asarray(dst_indices)
    T = np.asarray(T, dtype=np.float64)
    R, t = T[:3, :3], T[:3, 3]
    use = (idx >= 0) & (idx < len(d))
    p = s @ R.T + t
    e = np.zeros((len(s), 3))
    delta = p[use] - d[idx[use]]
    C = Cd[idx[use]] + R @ Cs[use] @ R.T
    C = 0.5 * (C + np.swapaxes(C, 1, 2))
    lam, V = np.linalg.eigh(C)
    W = np.einsum("nik,nk,njk->nij", V, 1.0 / np.sqrt(lam), V)
    ee = np.einsum("nij,nj->ni", W, delta)
    e[use] = ee
    sq = (ee * ee).sum(1)
    if huber > 0:
        out = sq > huber * huber
        w = np.where(out, huber / np.sqrt(np.maximum(sq, 1e-300)), 1.0)
        rho = np.where(out, 2.0 * huber * np.sqrt(sq) - huber * huber, sq)
    else:
        w, rho = np.ones_like(sq), sq
    pp = p[use]
    negpx = np.zeros((len(pp), 3, 3))
    negpx[:, 0, 1], negpx[:, 0, 2] = pp[:, 2], -pp[:, 1]
    negpx[:, 1, 0], negpx[:, 1, 2] = -pp[:, 2], pp[:, 0]
    negpx[:, 2, 0], negpx[:, 2, 1] = pp[:, 1], -pp[:, 0]
    J = np.concatenate([W @ negpx, W], axis=2)                      # n x 3 x 6
    A = np.einsum("n,nki,nkj->ij", w, J, J)
    b = np.einsum("n,nki,nk->i", w, J, ee)
    return dict(residuals=e, cost=0.5 * rho.sum(), A=A[np.triu_indices(6)], b=b, count=int(use.sum()), J=J, w=w)


def gicp_minimize(src, dst, src_covs, dst_covs, dst_indices, T0=None, max_iters=32, huber=0.5):
    """The 7-argument ComputeAlignment (align_gicp.cpp:41-117) with Levenberg-Marquardt on the Gauss-Newton normal
    equations of gicp_evaluate standing where the reference runs Ceres (absent): damping 1e-4 on diag A, / 3 after an
    accepted step, x 10 after a rejected one, the left-multiplied SE(3) update of pose_update. float64 throughout.
    Returns (pose, final cost, costs of the accepted poses)."""
    T = np.eye(4) if T0 is None else np.asarray(T0, dtype=np.float64).copy()
    iu = np.triu_indices(6)

    def normal_eq(r):
        A = np.zeros((6, 6)); A[iu] = r["A"]
        return A + A.T - np.diag(np.diag(A)), r["b"]

    good = gicp_evaluate(src, dst, src_covs, dst_covs, dst_indices, T, huber)
    lam, hist = 1e-4, [good["cost"]]
    for _ in range(max_iters):
        A, b = normal_eq(good)
        xi = np.linalg.solve(A + lam * np.diag(np.diag(A)), -b)
        trial = pose_update(xi, T)
        r = gicp_evaluate(src, dst, src_covs, dst_covs, dst_indices, trial, huber)
        if r["cost"] <= good["cost"]:
            T, good, lam = trial, r, max(lam / 3.0, 1e-9)
            hist.append(r["cost"])
        else:
            lam = min(lam * 10.0, 1e6)
    return T, good["cost"], hist


def covariances(pts, use_gicp=False, k=32):
    """ComputeCovariances (point_cloud_utils.cpp:100-161), float64 numpy restatement over scipy's exact k-NN."""
    from scipy.spatial import cKDTree
    p = np.asarray(pts, dtype=np.float64)
    _, idx = cKDTree(p).query(p, k=k + 1)
    nb = p[idx[:, 1:]]
    dl = nb - nb.mean(1, keepdims=True)
    C = np.einsum("nki,nkj->nij", dl, dl)
    if not use_gicp:
        return C / (k - 1)
    lam, V = np.linalg.eigh(C)                                       # ascending: column 0 = plane normal
    v = np.array([1e-2, 1.0, 1.0])
    return np.einsum("nik,k,njk->nij", V, v, V)
