"""Seeded synthetic RGB-D frames with known rigid motion (stands in for the camera;
the reference's only camera-free source is RandomSource, data_source.hpp:22-41).

Workload definitions follow SURVEY.md §8(d): D435-like intrinsics, depth scale
0.001 m/LSB, analytic room + spheres + box.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N

DEPTH_SCALE = 0.001

# nominal intrinsics per resolution (the reference hard-codes none)
INTRINSICS = {
    (640, 480): (385.0, 385.0, 320.0, 240.0),
    (848, 480): (424.0, 424.0, 424.0, 240.0),
    (1280, 720): (640.0, 640.0, 640.0, 360.0),
}


def intrinsics_for(w: int, h: int):
    if (w, h) in INTRINSICS:
        return INTRINSICS[(w, h)]
    f = 0.6 * w
    return (f, f, w / 2.0, h / 2.0)


def rot_xyz(rx: float, ry: float, rz: float) -> np.ndarray:
    """R_x(rx) @ R_y(ry) @ R_z(rz) — the composition of the reference's disabled
    self-test (rs_align_app.cpp:257-263: AngleAxis(0.1,x)*AngleAxis(-0.2,y)*AngleAxis(0.25,z))."""
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def make_pose(R: np.ndarray, t) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def rotvec_to_R(w) -> np.ndarray:
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * (K @ K)


def pose_error(T_est: np.ndarray, T_gt: np.ndarray):
    """(translation error in m, rotation error in rad) between two 4x4 poses."""
    D = np.linalg.inv(T_gt) @ T_est
    c = np.clip((np.trace(D[:3, :3]) - 1) / 2, -1, 1)
    return float(np.linalg.norm(D[:3, 3])), float(np.arccos(c))


@dataclass
class Noise:
    sigma_lsb_at_1m: float = 0.0
    p_invalid_pixel: float = 0.0
    p_invalid_block: float = 0.0


class Scene:
    def __init__(self, seed: int = 0):
        self._s = N.SynthScene()
        N.synth_lib().rst_synth_scene_default(seed, C.byref(self._s))

    def render(self, T_wc: np.ndarray, w: int, h: int, intr=None, noise: Noise | None = None,
               frame_seed: int = 0, rgb: bool = False, out: np.ndarray | None = None):
        """Renders depth (uint16 [h,w]) and optionally rgb (uint8 [h,w,3]) at camera->world pose T_wc."""
        fx, fy, cx, cy = intr if intr is not None else intrinsics_for(w, h)
        depth = out if out is not None else np.empty((h, w), dtype=np.uint16)
        assert depth.dtype == np.uint16 and depth.shape == (h, w) and depth.strides[1] == 2
        color = np.empty((h, w, 3), dtype=np.uint8) if rgb else None
        Tcm = np.ascontiguousarray(np.asarray(T_wc, dtype=np.float64).T)  # column-major
        nz = None
        if noise is not None:
            nz = N.SynthNoise(noise.sigma_lsb_at_1m, noise.p_invalid_pixel, noise.p_invalid_block)
        N.synth_lib().rst_synth_render(
            C.byref(self._s), Tcm.ctypes.data, fx, fy, cx, cy, w, h, DEPTH_SCALE,
            C.byref(nz) if nz is not None else None, frame_seed, depth.ctypes.data,
            depth.strides[0] // 2, color.ctypes.data if rgb else None)
        return (depth, color) if rgb else depth


def trajectory(n_frames: int, seed: int = 0, step_t: float = 0.015, step_r: float = 0.012) -> np.ndarray:
    """Smooth camera trajectory (camera->world poses, [n,4,4]): per-frame motion about
    step_t metres and step_r radians (~1-2 cm, ~0.5-1 deg)."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, size=6)
    T = np.empty((n_frames, 4, 4))
    for k in range(n_frames):
        s = k * 0.05
        # amplitudes chosen so that the per-frame derivative is ~step_t / step_r
        t = np.array([np.sin(s + ph[0]), 0.4 * np.sin(1.3 * s + ph[1]), 0.7 * np.sin(0.8 * s + ph[2])]) * (step_t / 0.05)
        r = np.array([0.5 * np.sin(0.9 * s + ph[3]), np.sin(1.1 * s + ph[4]), 0.6 * np.sin(0.7 * s + ph[5])]) * (step_r / 0.05)
        T[k] = make_pose(rotvec_to_R(r), t * 1.0)
    return T


def relative_pose(T_wc_dst: np.ndarray, T_wc_src: np.ndarray) -> np.ndarray:
    """Ground-truth T mapping src-camera points into the dst camera: p_dst = T p_src
    (the reference's convention, align_icp.cpp:107)."""
    return np.linalg.inv(T_wc_dst) @ T_wc_src


def render_sequence(n_frames: int, w: int, h: int, seed: int = 0, noise: Noise | None = None,
                    step_t: float = 0.015, step_r: float = 0.012, pinned=None):
    """Frames [n,h,w] uint16 + ground-truth frame-to-frame poses T_{k <- k+1} [n-1,4,4]."""
    scene = Scene(seed)
    Twc = trajectory(n_frames, seed, step_t, step_r)
    frames = pinned if pinned is not None else np.empty((n_frames, h, w), dtype=np.uint16)
    for k in range(n_frames):
        scene.render(Twc[k], w, h, noise=noise, frame_seed=seed * 100003 + k, out=frames[k])
    gt = np.stack([relative_pose(Twc[k], Twc[k + 1]) for k in range(n_frames - 1)])
    return frames, gt


def render_pairs(n_pairs: int, w: int, h: int, seed: int = 0, max_t: float = 0.03, max_r: float = np.deg2rad(2.0),
                 noise: Noise | None = None, first: int = 0, out=None):
    """Independent pairs with per-pair random motion: returns
    src [n,h,w], dst [n,h,w], gt [n,4,4] with p_dst = gt p_src. Pair k of the call is pair `first + k` of the
    seed's global list (a rank renders exactly its block of a sharded batch); `out` = (src, dst) buffers to fill."""
    src = out[0] if out is not None else np.empty((n_pairs, h, w), dtype=np.uint16)
    dst = out[1] if out is not None else np.empty((n_pairs, h, w), dtype=np.uint16)
    gt = np.empty((n_pairs, 4, 4))
    for k in range(n_pairs):
        i = first + k
        rng = np.random.default_rng(seed * 7919 + i)
        scene = Scene(seed * 7919 + i + 1)
        base = make_pose(rotvec_to_R(rng.normal(size=3) * 0.05), rng.uniform(-0.3, 0.3, size=3))
        d = rng.normal(size=3); d *= rng.uniform(0.3, 1.0) * max_t / np.linalg.norm(d)
        a = rng.normal(size=3); a *= rng.uniform(0.3, 1.0) * max_r / np.linalg.norm(a)
        T_dst = base
        T_src = base @ make_pose(rotvec_to_R(a), d)
        scene.render(T_dst, w, h, noise=noise, frame_seed=2 * i, out=dst[k])
        scene.render(T_src, w, h, noise=noise, frame_seed=2 * i + 1, out=src[k])
        gt[k] = relative_pose(T_dst, T_src)
    return src, dst, gt
