#!/usr/bin/env python
"""bench.py — frame-to-frame ICP throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the alignment hot path over one batch of synthetic input: a
640x480 depth sequence of F frames (F-1 frame-to-frame pairs, identity prior, 3-level
pyramid, (10,5,4) iterations) per GPU.  Per-GPU work is fixed (weak scaling); pairs are
independent, so there is no collective on the data path — NCCL only gathers the per-pair
poses (one all_gather per step).

  value    : pairs/s with the depth frames already resident in HBM
  e2e      : pairs/s through the public call (rst_align_sequence) with PINNED HOST frames:
             H2D of all frames and D2H of poses + statistics inside the timed region
  roofline : level-0 fused ICP kernel, CUDA events around the 10 level-0 launches of each step
  cpu_baseline : Oracle-R (restatement of the reference's AlignIcp3d path) on the host cores

`--impl reference` times only that CPU path (the reference has no GPU code and cannot be
built here).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H = 640, 480
FRAMES = 129                    # 128 pairs per GPU per step
ITERS = (10, 5, 4)
SURVEY_BYTES_PER_PX = 36        # src vertex 12 + dst vertex 12 + dst normal 12
METRIC = "icp_frame_pairs_per_sec_640x480"


def cm_to_pose_np(p):
    return np.swapaxes(np.asarray(p).reshape(-1, 4, 4), -1, -2).astype(np.float64)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the level-0 ICP kernel from the committed ncu capture, if any."""
    p = ROOT / "profiles" / "icp_l0_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return None
    return None


def bind_to_gpu_numa(props, local_rank: int = 0, local_world: int = 1) -> str:
    """Best effort: run this rank on the cores of the NUMA node its GPU hangs off, so that the pinned
    frame buffers (first touch) and the H2D copies stay on the local socket. With 8 ranks the host memory
    system, not PCIe, limits the end-to-end figure otherwise. When the platform does not expose the GPU's NUMA
    node (virtualised hosts report -1), the ranks at least take DISJOINT slices of the allowed cores, so that the
    upload threads of different ranks never share a core."""
    def spread() -> str:
        cpus = sorted(os.sched_getaffinity(0))
        if local_world <= 1 or len(cpus) < local_world:
            return "numa node unknown: not bound"
        per = len(cpus) // local_world
        mine = cpus[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, set(mine))
        return f"numa node unknown: cores {mine[0]}-{mine[-1]} ({len(mine)} of {len(cpus)}, disjoint per rank)"
    try:
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node < 0:
            return spread()
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"numa node {node}: no allowed cpu"
        os.sched_setaffinity(0, cpus)
        return f"numa node {node}, {len(cpus)} cpus"
    except Exception as e:  # not fatal: the bench still runs, only possibly across sockets
        try:
            return spread() + f" ({type(e).__name__} reading the NUMA node)"
        except Exception:
            return f"not bound ({type(e).__name__})"


class ClockSampler:
    """SM clock / power / throttle reasons sampled while the timed region runs: NVML polled every 5 ms from a thread
    (the 50-step headline region lasts 0.1 s), `nvidia-smi -lms 50` when pynvml is not importable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, uuid: str | None, index: int = 0):
        self.rows, self.proc, self.t, self.nvml, self.stop_flag = [], None, None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid) if uuid else pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nvml, self.h = pynvml, h
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        cmd = ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"]
        if uuid:
            cmd += ["-i", uuid]
        try:
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)), "hw_slowdown"),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)), "hw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)), "sw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)), "sw_power_cap")]
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        mx = float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM))
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                r = int(get_reasons(self.h))
                self.rows.append((time.time(), (sm, mx, pw, [nm for b, nm in bits if r & b])))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 7:
                continue
            try:
                row = (float(f[0]), float(f[1]), float(f[2]), [nm for nm, val in zip(self.NAMES, f[3:7]) if val.lower().startswith("active")])
            except ValueError:
                continue
            self.rows.append((time.time(), row))

    def stop(self, t0: float, t1: float) -> dict:
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sel = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (t, r) in self.rows if t0 <= t <= t1 + 0.05] or [r for (_, r) in self.rows]
        sm = [r[0] for r in sel]; mx = [r[1] for r in sel]; pw = [r[2] for r in sel]
        reasons = sorted({nm for r in sel for nm in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons,
                "source": "nvml, 5 ms" if self.nvml is not None else "nvidia-smi -lms 50"}


# --------------------------------------------------------------------------------------------
# CPU arm: Oracle-R (the reference's own algorithm) on the host cores
# --------------------------------------------------------------------------------------------
def cpu_reference_kind():
    """The CPU arm times Oracle-R, the restatement ("port"). The reference's own align_icp.cpp +
    point_cloud_utils.cpp do compile here against stand-in headers (oracle/_ref) and Oracle-R reproduces
    that build bit for bit (tests/test_reference_compiled.py), but the stand-in Eigen/nanoflann make it
    ~5x slower than the restatement, so timing it would flatter the GPU: the faster one is the baseline."""
    return "port"


def cpu_reference_rate(frames: np.ndarray, intr, n_pairs: int, threads: int):
    """pairs/s of the reference's CPU path (Oracle-R) over `n_pairs` consecutive frame pairs with `threads` threads."""
    from oracle import oracle as O
    src = np.ascontiguousarray(frames[1:n_pairs + 1])
    dst = np.ascontiguousarray(frames[0:n_pairs])
    t0 = time.perf_counter()
    ok, T = O.align_depth_pairs(src, dst, intr, voxel=0.05, max_iter=128, n_threads=threads)
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt, T, ok


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from realsensetracker_b200 import synth
    cores = host_cores()
    n_pairs = max(cores, 2)                      # one pair per host thread per step
    intr = synth.intrinsics_for(W, H)
    frames, gt = synth.render_sequence(n_pairs + 1, W, H, seed=0)
    for _ in range(args.warmup):
        cpu_reference_rate(frames, intr, min(n_pairs, cores), cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, T, ok = cpu_reference_rate(frames, intr, n_pairs, cores)
    dt = time.perf_counter() - t0
    value = n_pairs * args.steps / dt
    errs = [synth.pose_error(T[i], gt[i]) for i in range(n_pairs)]
    sample = (f"{n_pairs} consecutive frame pairs of the 640x480 sequence per step, one pair per thread: "
              "back-project -> RemoveNans -> DownsampleVoxel(0.05) -> AlignIcp3d(128 it, leaf 16); "
              "g++ -O2 -ffp-contract=off (the reference sets no flags); Oracle-R restatement, bit-identical to the "
              "reference's own align_icp.cpp compiled against stand-in headers (oracle/_ref), and ~5x faster than that build")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"seq{W}x{H}_f2f_3level_icp" + (" (BASELINE.json configs[1])" if (W, H) == (640, 480) else " (ad-hoc size)"),
                   "width": W, "height": H,
                   "sample": "a bounded sample of that workload per step: the first %d frames (%d frame pairs) of the same synthetic sequence" % (n_pairs + 1, n_pairs),
                   "frames_per_step": n_pairs + 1, "pairs_per_step": n_pairs,
                   "algorithm": "the reference's own alignment of a frame pair: AlignIcp3d (KD-tree point-to-point, GNC Geman-McClure, Kabsch; "
                                "one level, 128 iterations) on 5 cm voxel clouds, CPU"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": cpu_reference_kind(), "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "pose_err_vs_gt": {"t_m_max": max(e[0] for e in errs), "r_rad_max": max(e[1] for e in errs)},
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_gpu(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from realsensetracker_b200 import Aligner, default_params, synth
    from realsensetracker_b200 import _native as N

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the alignment path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(torch.cuda.get_device_properties(local_rank), local_rank, env_int("LOCAL_WORLD_SIZE", world)) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    intr = synth.intrinsics_for(W, H)
    P = default_params(num_levels=3, iters=list(ITERS))
    n_pairs = FRAMES - 1

    # synthetic sequence of this rank, rendered straight into pinned host memory
    pinned = torch.empty((FRAMES, H, W), dtype=torch.int16, pin_memory=True)
    frames = pinned.numpy().view(np.uint16)
    _, gt = synth.render_sequence(FRAMES, W, H, seed=rank, pinned=frames)
    d_frames = pinned.to(dev)                                   # HBM-resident copy for `value`

    stream = torch.cuda.Stream(device=dev)
    al = Aligner(W, H, FRAMES, n_pairs, device=local_rank, stream=stream.cuda_stream)
    if args.chunk is not None:
        al.set_pipeline_chunk(args.chunk)
    if args.no_split:
        al.set_stream_split(0)
    src_slots = np.arange(1, FRAMES, dtype=np.int32)
    dst_slots = np.arange(0, FRAMES - 1, dtype=np.int32)
    d_poses = torch.empty((n_pairs, 16), dtype=torch.float32, device=dev)
    d_all = torch.empty((world * n_pairs, 16), dtype=torch.float32, device=dev) if world > 1 else None
    h_poses = np.tile(np.eye(4, dtype=np.float32).reshape(16), (n_pairs, 1))
    h_stats = (N.Stats * n_pairs)()
    K = N.Intrinsics(*intr)
    from realsensetracker_b200.align import _frames
    fr = _frames(frames)
    lib, ctx = al._lib, al._ctx

    def step_resident():
        al.begin(W, H, intr, P)   # P is rebound for the early-exit region below
        al.set_frames_device(d_frames.data_ptr(), FRAMES, W, W * H)
        al.preprocess(0, FRAMES)
        al.align_slots(src_slots, dst_slots, fetch=False)
        al.copy_results_device(d_poses.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_poses)

    # e2e: two contexts driven alternately (submit step k, then wait for step k-1), so the PCIe copies
    # of one step overlap the kernels of the other. Every step still uploads all of its frames from
    # pinned host memory and reads all of its poses + statistics back to the host.
    stream_b = torch.cuda.Stream(device=dev)
    al_b = Aligner(W, H, FRAMES, n_pairs, device=local_rank, stream=stream_b.cuda_stream)
    ctxs = [ctx, al_b._ctx]

    def e2e_steps(steps):
        pending = None
        for k in range(steps):
            cur = ctxs[k & 1]
            rc = lib.rst_align_sequence_async(cur, fr, FRAMES, C.byref(K), C.byref(P), None)
            if rc != 0:
                raise RuntimeError(lib.rst_last_error(cur).decode())
            if pending is not None and lib.rst_wait(pending, h_poses.ctypes.data, C.addressof(h_stats)) != 0:
                raise RuntimeError(lib.rst_last_error(pending).decode())
            pending = cur
        if pending is not None and lib.rst_wait(pending, h_poses.ctypes.data, C.addressof(h_stats)) != 0:
            raise RuntimeError(lib.rst_last_error(pending).decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup):
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step_fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.time()
            l0 = al.launch_count
            e0.record(stream)
            for _ in range(steps):
                step_fn()
            e1.record(stream)
            barrier()
            t1 = time.time()
            ms = e0.elapsed_time(e1)
            launches = al.launch_count - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, t0, t1

    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        pass
    sampler = ClockSampler(uuid) if rank == 0 else None

    # ---- value: HBM-resident inputs, the library's default schedule (two-stream split of the batch)
    for _ in range(args.warmup):
        with torch.cuda.stream(stream):
            step_resident()
    ms_val, launches, t0, t1 = timed(step_resident, args.steps, 0)
    clocks = sampler.stop(t0, t1) if sampler else None
    total_pairs = world * n_pairs * args.steps
    value = total_pairs / (ms_val * 1e-3)

    # ---- the same steps with the per-level convergence test on (converge_eps): reported separately, the
    #      headline keeps the fixed iteration count so that the work per pair is constant
    P_fixed = P
    P = default_params(num_levels=3, iters=list(ITERS), converge_eps=2e-5)
    with torch.cuda.stream(stream):
        for _ in range(2):
            step_resident()
    ms_early, _, _, _ = timed(step_resident, args.steps, 0)
    early_poses = cm_to_pose_np(d_poses.cpu().numpy())
    early_err = np.array([synth.pose_error(early_poses[i], gt[i]) for i in range(n_pairs)])
    P = P_fixed

    # ---- roofline region: the same steps with CUDA events around every kernel group (the fused launch; one
    #      pre-processing launch per level), all on ONE stream so that each is timed in isolation
    al.set_stream_split(0)
    with torch.cuda.stream(stream):
        for _ in range(2):
            step_resident()
    al.profile_enable(True)
    ms_single, _, _, _ = timed(step_resident, args.steps, 0)
    prof = al.profile_read()
    al.profile_enable(False)

    # ---- the other two schedules, for comparison (rst_set_schedule): fully fused (ONE cluster launch for all 19
    #      iterations of a batch) and hybrid (coarse levels fused, finest level one launch per iteration)
    def schedule_region(sched):
        al.set_schedule(sched)
        al.set_stream_split(0)
        with torch.cuda.stream(stream):
            for _ in range(2):
                step_resident()
        al.profile_enable(True)
        ms_1s, _, _, _ = timed(step_resident, args.steps, 0)
        pr = al.profile_read()
        al.profile_enable(False)
        al.set_stream_split(0 if args.no_split else 32)
        ms_sp, _, _, _ = timed(step_resident, args.steps, 0)
        return ms_1s, ms_sp, pr
    ms_fused_1s, ms_fused, prof_fused = schedule_region(1)
    ms_hybrid_1s, ms_hybrid, prof_hy = schedule_region(3)
    al.set_schedule(0)
    with torch.cuda.stream(stream):
        step_resident()     # d_poses holds the default schedule's result again
    torch.cuda.synchronize()

    # correctness of what was timed: poses vs ground truth, and (N>1) the gathered block of this rank
    poses_dev = d_poses.cpu().numpy()
    from realsensetracker_b200.align import cm_to_pose
    Tres = cm_to_pose(poses_dev)
    errs = np.array([synth.pose_error(Tres[i], gt[i]) for i in range(n_pairs)])
    if world > 1:
        mine = d_all[rank * n_pairs:(rank + 1) * n_pairs].cpu().numpy()
        assert np.array_equal(mine, poses_dev), "all_gather block differs from the local result"

    # ---- e2e: pinned host frames in, poses + stats out, every step
    e2e_steps(args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = al.launch_count + al_b.launch_count
    e0.record(stream)
    e2e_steps(args.steps)
    e1.record(stream)          # both contexts have been waited for: the device is idle here
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    launches_e2e = al.launch_count + al_b.launch_count - l0
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = total_pairs / (ms_e2e * 1e-3)

    # the plain blocking call (one context, nothing overlapped), for reference beside the streaming form
    def step_blocking():
        h_poses[:] = np.eye(4, dtype=np.float32).reshape(16)   # poses are in/out: every step starts from the identity prior
        rc = lib.rst_align_sequence(ctx, fr, FRAMES, C.byref(K), C.byref(P), h_poses.ctypes.data, C.addressof(h_stats))
        if rc != 0:
            raise RuntimeError(lib.rst_last_error(ctx).decode())
    ms_blk, _, _, _ = timed(step_blocking, max(5, args.steps // 2), 2)
    blocking_value = world * n_pairs * max(5, args.steps // 2) / (ms_blk * 1e-3)
    Te2e = cm_to_pose(h_poses)
    assert np.array_equal(Te2e, Tres), "host-path and device-resident results differ"
    status_bad = sum(1 for s in h_stats if s.status != 0)
    mean_count = float(np.mean([s.count for s in h_stats]))

    # ---- sustained: the same resident steps for >= 3 s with the clocks sampled inside (>= 20 samples)
    sustained = None
    if not args.quick:
        n_sus = max(int(3.2e3 / (ms_val / args.steps)), args.steps)
        sus_sampler = ClockSampler(uuid) if rank == 0 else None
        ms_sus, _, ts0, ts1 = timed(step_resident, n_sus, 0)
        sus_clocks = sus_sampler.stop(ts0, ts1) if sus_sampler else None
        sustained = {"value": world * n_pairs * n_sus / (ms_sus * 1e-3), "unit": "pairs/s", "steps": n_sus, "seconds": ms_sus * 1e-3,
                     "ms_per_step": ms_sus / n_sus, "clocks": sus_clocks}

    # ---- single-pair latency: one blocking rst_align_pairs call per pair (the reference's own use: one pair at a
    #      time, rs_replay_app.cpp:211-287), host frames in, pose out; both tilings
    latency = None
    if rank == 0 and world == 1 and not args.quick:
        latency = {}
        al1 = Aligner(W, H, 2, 1, device=local_rank)
        for name, til in (("throughput_tiling", 0), ("latency_tiling", 1)):
            P1 = default_params(num_levels=3, iters=list(ITERS), tiling=til)
            two = frames[:2]
            for _ in range(20):
                T1, st1 = al1.align_sequence(two, intr, P1)
            t0l = time.perf_counter()
            nl = 200
            for _ in range(nl):
                T1, st1 = al1.align_sequence(two, intr, P1)
            dtl = (time.perf_counter() - t0l) / nl
            latency[name] = {"blocking_pair_us": dtl * 1e6, "pose_err_vs_gt_t_m": float(synth.pose_error(T1[0], gt[0])[0])}
        al1.close()
        latency["what"] = "wall clock of one blocking rst_align_sequence call on 2 pinned host frames (H2D + 3 pre-processing + %d iteration launches + D2H), mean of 200" % sum(ITERS)

    # ---- concurrent pinned H2D of all ranks (the ceiling of the end-to-end figure at N > 1)
    h2d_ceiling = None
    if not args.quick:
        d_tmp = torch.empty_like(d_frames)
        hs = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(hs):
            d_tmp.copy_(pinned, non_blocking=True)
            hs.synchronize()
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(hs)
            for _ in range(10):
                d_tmp.copy_(pinned, non_blocking=True)
            ev1.record(hs)
            hs.synchronize()
        gbs = 10 * pinned.numel() * 2 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
        if world > 1:
            tt = torch.tensor([gbs], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MIN)
            gbs = float(tt.item())
        del d_tmp
        h2d_ceiling = {"gbs_per_gpu_min_over_ranks": gbs, "what": "10 copies of the step's %d MB pinned frame buffer, all %d rank(s) at the same time" % (pinned.numel() * 2 // 1000000, world),
               "e2e_ceiling_pairs_per_s": world * n_pairs / (pinned.numel() * 2 / (gbs * 1e9))}

    # ---- the other BASELINE.json configurations (extra keys; the headline stays configs[1])
    configs_extra = None
    if not args.quick:
        configs_extra = run_extra_configs(args, rank, world, local_rank, dev, stream, measured_peak()[0])

    if rank == 0:
        peak, peak_src = measured_peak()
        npx = W * H
        assoc_frac = mean_count / npx                                  # associated fraction of the finest level (last iterate)
        # compulsory bytes of this layout per pair-iteration on level l: 2 B x pixels + 16 B x associated pixels
        lvl_px = [(W >> l) * (H >> l) for l in range(3)]
        bytes_pair_iter = [2.0 * p + 16.0 * assoc_frac * p for p in lvl_px]
        n_l0 = max(prof.launches_icp[0], 1)
        t_l0 = prof.ms_icp[0] * 1e-3 / n_l0                            # average level-0 launch, seconds
        bytes_layout = n_pairs * bytes_pair_iter[0]                    # per level-0 launch
        bytes_survey = n_pairs * float(SURVEY_BYTES_PER_PX) * npx
        ach = bytes_layout / t_l0 / 1e9
        ach_s = bytes_survey / t_l0 / 1e9
        t_coarse = prof_hy.ms_icp_fused * 1e-3 / max(prof_hy.launches_icp_fused, 1)
        bytes_coarse = n_pairs * sum(ITERS[l] * bytes_pair_iter[l] for l in (1, 2))
        lvl_us = [prof.ms_icp[l] * 1e3 / max(prof.launches_icp[l], 1) for l in range(3)]
        t_ff = prof_fused.ms_icp_fused * 1e-3 / max(prof_fused.launches_icp_fused, 1)
        bytes_all = n_pairs * sum(ITERS[l] * bytes_pair_iter[l] for l in range(3))
        tr = ncu_traffic()
        pre_ms = sum(prof.ms_preprocess[l] for l in range(3))
        roofline = {
            "kernel": "k_icp_iter<level 0> (fused association + point-to-plane J^T J / J^T r + reduction + solve), %d launches per step" % ITERS[0],
            "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": (tr or {}).get("dram_bytes_per_launch"),
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bytes_layout,
            "bytes_per_px": {"src_depth": 2, "dst_geometry_gather": 16, "associated_px_per_pair": mean_count},
            "avg_launch_us": t_l0 * 1e6, "launches_timed": prof.launches_icp[0],
            "survey_equiv": {"bytes_per_px": SURVEY_BYTES_PER_PX, "achieved": ach_s, "frac": ach_s / peak},
            "level_launch_us": lvl_us,
            "level_frac": [n_pairs * bytes_pair_iter[l] / (lvl_us[l] * 1e-6) / 1e9 / peak if lvl_us[l] > 0 else None for l in range(3)],
            "stage_share_of_step": {"icp_l0": prof.ms_icp[0] / ms_single, "icp_l1": prof.ms_icp[1] / ms_single,
                                    "icp_l2": prof.ms_icp[2] / ms_single, "preprocess": pre_ms / ms_single,
                                    "other": max(0.0, 1.0 - (sum(prof.ms_icp[l] for l in range(3)) + pre_ms) / ms_single)},
            "schedules": {
                "auto": {"what": "default = one launch per iteration per level (k_icp_iter), two-stream split for large batches", "value": value},
                "fused": {"what": "rst_set_schedule(1): k_icp_fused, one cluster of CTAs per pair, all %d iterations of all levels in ONE launch "
                                  "(DSMEM reduction, no global partials)" % sum(ITERS),
                          "value": total_pairs / (ms_fused * 1e-3), "launch_us": t_ff * 1e6, "achieved": bytes_all / t_ff / 1e9,
                          "frac": bytes_all / t_ff / 1e9 / peak},
                "hybrid": {"what": "rst_set_schedule(3): levels 2 and 1 (%d + %d iterations) fused in one launch, level 0 one launch per iteration" % (ITERS[2], ITERS[1]),
                           "value": total_pairs / (ms_hybrid * 1e-3), "coarse_launch_us": t_coarse * 1e6,
                           "coarse_frac": bytes_coarse / t_coarse / 1e9 / peak if t_coarse > 0 else None}},
            "measured_in": "a second timed region of the same steps on one stream with CUDA events around every kernel group (ms_per_step %.3f)" % (ms_single / args.steps),
        }
        # CPU baseline (bounded sample) — rank 0, N=1 only
        cpu = None
        parity_ref = None
        if world == 1 and not args.no_cpu:
            cores = host_cores()
            n_cpu = max(8, min(2 * cores, n_pairs))
            rate, dt, Tc, ok = cpu_reference_rate(frames, intr, n_cpu, cores)
            cerr = np.array([synth.pose_error(Tc[i], gt[i]) for i in range(n_cpu)])
            diff = np.array([synth.pose_error(Tres[i], Tc[i]) for i in range(n_cpu)])
            parity_ref = {
                "what": "the SAME %d frame pairs through the reference's align path (CPU, AlignIcp3d on 5 cm voxel clouds) and through this "
                        "repo's CUDA path, both scored against the known motion" % n_cpu,
                "ours_err_vs_gt": {"t_m_max": float(errs[:n_cpu, 0].max()), "r_rad_max": float(errs[:n_cpu, 1].max())},
                "reference_err_vs_gt": {"t_m_max": float(cerr[:, 0].max()), "r_rad_max": float(cerr[:, 1].max())},
                "ours_vs_reference_pose": {"t_m_max": float(diff[:, 0].max()), "r_rad_max": float(diff[:, 1].max())},
                "ours_not_worse_on_every_pair": bool(np.all(errs[:n_cpu, 0] <= cerr[:, 0] + 1e-4) and np.all(errs[:n_cpu, 1] <= cerr[:, 1] + 1e-4)),
                "note": "different algorithms by the north star's definition: the pose difference is the reference's own error, not ours "
                        "(tests/test_vs_reference.py asserts ours <= reference against ground truth, with the reference compiled from its own source)"}
            rate1, _, _, _ = cpu_reference_rate(frames, intr, 3, 1)    # as the reference itself runs: one thread
            from oracle import oracle as O
            t0n = time.perf_counter()
            O.align_pair(frames[1], frames[0], intr, O.default_params())
            dtn = time.perf_counter() - t0n
            cpu = {"value": rate, "unit": "pairs/s", "cores": cores, "kind": cpu_reference_kind(),
                   "sample": f"reference AlignIcp3d path (voxel 0.05, 128 it; Oracle-R restatement, bit-identical to the reference's "
                             f"own align_icp.cpp compiled against stand-in headers) "
                             f"on the first {n_cpu} pairs of the same sequence, one pair per thread, {dt:.1f} s wall",
                   "pose_err_vs_gt": {"t_m_max": float(cerr[:, 0].max()), "r_rad_max": float(cerr[:, 1].max())},
                   "reference_path_1thread_pairs_per_s": rate1,
                   "same_algorithm_port_1thread_pairs_per_s": 1.0 / dtn}
        # the reference's OWN algorithm on the GPU (rst_icp3d_depth: back-project -> voxel 0.05 -> AlignIcp3d 128 it),
        # host frames in, poses out — the same-algorithm comparison for the CPU figure above
        ref_gpu = None
        if world == 1 and not args.no_cpu:
            sidx, didx = np.arange(1, FRAMES, dtype=np.int32), np.arange(0, FRAMES - 1, dtype=np.int32)
            al.icp3d_depth(frames, sidx, didx, intr)                       # warm-up (allocations)
            # wall clock of the blocking call AND, beside it, CUDA events on the context's stream around the same calls
            # (that stream waits for the chunked uploads, so the span covers H2D + depth -> cloud + ICP + D2H)
            evr0, evr1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0r = time.perf_counter()
            reps = 3
            evr0.record(stream)
            for _ in range(reps):
                okr, Tr, mcr, cnts = al.icp3d_depth(frames, sidx, didx, intr)
            evr1.record(stream)
            evr1.synchronize()
            dtr = (time.perf_counter() - t0r) / reps
            dev_ms_batch = evr0.elapsed_time(evr1) / reps
            rerr = np.array([synth.pose_error(Tr[i], gt[i]) for i in range(n_pairs)])
            searched, queried = al.icp3d_cache_stats()
            it_run, it_asked = al.icp3d_iteration_stats()
            al.set_icp3d_cache(0.0, 0.0, 0.0)                              # every point searches in every iteration ...
            al.set_icp3d_fixed_point_skip(False)                           # ... and every iteration runs: the reference's schedule as written
            t0n = time.perf_counter()
            okn, Tn_, _, _ = al.icp3d_depth(frames, sidx, didx, intr)
            dt_nocache = time.perf_counter() - t0n
            al.set_icp3d_cache()
            al.set_icp3d_fixed_point_skip(True)
            cache = {"searched": searched, "queried": queried, "searched_fraction": searched / max(queried, 1),
                     "iterations_run": it_run, "iterations_asked": it_asked,
                     "without_cache_and_skip_pairs_per_s": n_pairs / dt_nocache, "poses_bit_identical_to_without": bool(np.array_equal(Tn_, Tr)),
                     "what": "a query searches only when |p' - nbr| + (motion since the entry was proven) >= the proven radius (triangle inequality); "
                             "an iteration is not run when the previous one returned its pose bit for bit (it would repeat it until mu changes); "
                             "indices, distances and poses are those of searching every point in every iteration, bit for bit"}
            # one pair at a time, as the reference's caller runs it (rs_replay_app.cpp:246-251): one CTA per pair vs the
            # automatic thread-block cluster per pair
            single = {}
            for name, ctas in (("one_cta_per_pair", 1), ("cluster_per_pair", 0)):
                al.set_icp3d_cluster(ctas)
                al.icp3d_depth(frames[:2], sidx[:1], didx[:1], intr)
                t0s = time.perf_counter()
                evr0.record(stream)
                for _ in range(10):
                    al.icp3d_depth(frames[:2], sidx[:1], didx[:1], intr)
                evr1.record(stream)
                evr1.synchronize()
                single[name + "_ms"] = (time.perf_counter() - t0s) / 10 * 1e3
                single[name + "_cuda_events_ms"] = evr0.elapsed_time(evr1) / 10
            al.set_icp3d_cluster(0)
            # the cloud utilities and the GICP alignment of the reference's align library on one pair's clouds
            # (host clouds in, results out, best of 5 blocking calls each)
            src_c, dst_c = al.icp3d_read_cloud(1, int(cnts[1])), al.icp3d_read_cloud(0, int(cnts[0]))

            def best_ms(f, reps=5):
                f()
                best = 1e9
                for _ in range(reps):
                    t0u = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0u)
                return best * 1e3
            utils = {"points": [int(len(src_c)), int(len(dst_c))],
                     "find_correspondences_ms": best_ms(lambda: al.find_correspondences(dst_c, src_c)),
                     "compute_normals_k16_ms": best_ms(lambda: al.cloud_normals(src_c, k=16)),
                     "compute_covariances_ms": best_ms(lambda: al.cloud_covariances(src_c)),
                     "downsample_voxel_ms": best_ms(lambda: al.downsample_voxel(src_c, 0.1)),
                     "align_icp3d_128it_ms": best_ms(lambda: al.icp3d_pairs([src_c], [dst_c], 128)),
                     "gicp_align_16x4_ms": best_ms(lambda: al.gicp_align(src_c, dst_c))}
            Tg_, _ = al.gicp_align(src_c, dst_c)
            utils["gicp_pose_err_vs_gt_t_m"] = float(synth.pose_error(Tg_, gt[0])[0])
            ref_gpu = {"value": n_pairs / dtr, "single_pair": single, "unit": "pairs/s", "ms_per_step": dtr * 1e3, "timing": "host wall clock, H2D + D2H included",
                       "cuda_events_ms_per_step": dev_ms_batch, "cuda_events_pairs_per_s": n_pairs / (dev_ms_batch * 1e-3),
                       "algorithm": "reference AlignIcp3d on the GPU: exact grid NN, GM/GNC weights, Kabsch, 128 iterations, voxel 0.05",
                       "mean_cloud_points": float(np.mean(cnts)), "pairs_ok": int(okr.sum()), "neighbour_cache": cache, "cloud_utilities": utils,
                       "pose_err_vs_gt": {"t_m_max": float(rerr[:, 0].max()), "r_rad_max": float(rerr[:, 1].max())}}
        h2d = FRAMES * H * W * 2 + n_pairs * (8 + 64)
        d2h = n_pairs * (64 + C.sizeof(N.Stats))
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_val / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"seq{W}x{H}_f2f_3level_icp" + (" (BASELINE.json configs[1])" if (W, H) == (640, 480) else " (ad-hoc size)"), "width": W, "height": H,
                       "frames_per_gpu_per_step": FRAMES, "pairs_per_gpu_per_step": n_pairs, "levels": 3,
                       "iters_fine_to_coarse": list(ITERS), "accumulation": "fp32 per block (<=8192 px), fp64 across blocks and in the solve",
                       "l2_policy": f"inputs larger than L2: {FRAMES * npx * (2 + 16) * 1.3125 / 1e6:.0f} MB of depth pyramid + geometry maps streamed per step (L2 = 126 MB)",
                       "parallelism": f"pairs sharded, {world} rank(s), NCCL all_gather of poses only", "host_binding_rank0": numa},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": ms_e2e / args.steps, "gpu_launches": int(launches_e2e),
                    "api": "rst_align_sequence_async + rst_wait, two contexts per GPU alternating (copy of step k+1 under the kernels of step k)",
                    "h2d_ceiling": h2d_ceiling},
            "blocking_call": {"value": blocking_value, "unit": "pairs/s", "ms_per_step": ms_blk / max(5, args.steps // 2),
                              "api": "rst_align_sequence: ONE context, one blocking call per step — what a caller written like the reference's loop "
                                     "gets (the call runs the batch as two halves, the upload of the second under the kernels of the first)"},
            "sustained": sustained,
            "latency": latency,
            "configs": configs_extra,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity_vs_reference": parity_ref,
            "reference_algorithm_on_gpu": ref_gpu,
            "early_exit": {"value": total_pairs / (ms_early * 1e-3), "unit": "pairs/s", "converge_eps": 2e-5,
                           "ms_per_step": ms_early / args.steps,
                           "pose_err_vs_gt": {"t_m_max": float(early_err[:, 0].max()), "r_rad_max": float(early_err[:, 1].max())}},
            "pose_err_vs_gt": {"t_m_max": float(errs[:, 0].max()), "r_rad_max": float(errs[:, 1].max()),
                               "pairs_failed": status_bad},
        }
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    al.close()
    al_b.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# The other BASELINE.json configurations: c3 (1280x720, 256 pairs, strong-scaled), c4 (848x480 RGB-D,
# frame-to-keyframe, geometric + photometric), c5 (1280x720 stress: 30 % invalid depth, up to 10 cm / 8 deg,
# Huber). Each: pairs/s with the frames resident in HBM, its own level-0 launch time against the measured HBM
# peak, pose error against the known motion.
# --------------------------------------------------------------------------------------------
def run_extra_configs(args, rank, world, local_rank, dev, stream, peak):
    import torch
    import torch.distributed as dist
    from realsensetracker_b200 import Aligner, default_params, shard, synth
    from realsensetracker_b200 import _native as N
    out = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_steps(step_fn, steps):
        with torch.cuda.stream(stream):
            for _ in range(3):
                step_fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                step_fn()
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def reduce_max(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def pair_batch(key, w, h, total, P, render_kw, extra_bytes_px, what):
        """`total` independent pairs partitioned over the ranks (strong scaling): rank r renders and aligns pairs
        shard.partition(total, world, r); the per-pair poses are all-gathered like the headline's."""
        intr = synth.intrinsics_for(w, h)
        lo, hi = shard.partition(total, world, rank)
        n = hi - lo
        pin = torch.empty((2 * n, h, w), dtype=torch.int16, pin_memory=True)
        fr = pin.numpy().view(np.uint16)
        _, _, gt = synth.render_pairs(n, w, h, first=lo, out=(fr[n:], fr[:n]), **render_kw)   # dst in slots [0, n), src in [n, 2n)
        d_fr = pin.to(dev)
        al = Aligner(w, h, 2 * n, n, device=local_rank, stream=stream.cuda_stream)
        src_slots = np.arange(n, 2 * n, dtype=np.int32); dst_slots = np.arange(0, n, dtype=np.int32)
        d_poses = torch.empty((n, 16), dtype=torch.float32, device=dev)
        sizes = [shard.partition(total, world, r)[1] - shard.partition(total, world, r)[0] for r in range(world)]
        d_pad = torch.zeros((max(sizes), 16), dtype=torch.float32, device=dev)
        d_all = [torch.empty_like(d_pad) for _ in range(world)] if world > 1 else None
        ngate = P.normal_cos_min > -1.0

        def step():
            al.begin(w, h, intr, P)
            al.set_frames_device(d_fr.data_ptr(), 2 * n, w, w * h)
            al.preprocess(0, 2 * n)
            al.align_slots(src_slots, dst_slots, fetch=False)
            al.copy_results_device(d_poses.data_ptr())
            if world > 1:
                d_pad[:n].copy_(d_poses)
                dist.all_gather(d_all, d_pad)
        steps = max(3, min(args.steps, 10))
        ms = time_steps(step, steps)
        al.set_stream_split(0)
        al.profile_enable(True)
        time_steps(step, 3)
        prof = al.profile_read()
        al.profile_enable(False)
        torch.cuda.synchronize()
        from realsensetracker_b200.align import cm_to_pose
        T = cm_to_pose(d_poses.cpu().numpy())
        st = (N.Stats * n)()
        errs = np.array([synth.pose_error(T[i], gt[i]) for i in range(n)]) if n else np.zeros((0, 2))
        # statistics of the last iterate (associated pixels) from the device
        d_stats = torch.empty((n * C.sizeof(N.Stats),), dtype=torch.uint8, device=dev)
        with torch.cuda.stream(stream):
            al.copy_results_device(None, d_stats.data_ptr())
        torch.cuda.synchronize()
        h_stats_raw = d_stats.cpu().numpy()      # keep the host array alive across the memmove
        C.memmove(C.addressof(st), h_stats_raw.ctypes.data, n * C.sizeof(N.Stats))
        count = float(np.mean([s.count for s in st])) if n else 0.0
        failed = sum(1 for s in st if s.status != 0)
        t_l0 = prof.ms_icp[0] * 1e-3 / max(prof.launches_icp[0], 1)
        bytes_l0 = n * ((2.0 + (16.0 if ngate else 0.0)) * w * h + (16.0 + extra_bytes_px) * count)
        ach = bytes_l0 / t_l0 / 1e9 if t_l0 > 0 else 0.0
        e_t, e_r, nf, frac = reduce_max([errs[:, 0].max() if n else 0.0, errs[:, 1].max() if n else 0.0, failed, ach / peak])
        al.close()
        out[key] = {"what": what, "width": w, "height": h, "pairs_total": total, "pairs_this_rank": n, "scaling": "strong",
                    "value": total * steps / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms / steps,
                    "level0": {"launch_us": t_l0 * 1e6, "achieved_gbs": ach, "frac_of_measured_peak_max_over_ranks": frac,
                               "associated_px_per_pair": count, "bytes_per_px": "2 src depth%s + 16 gather%s" % (" + 16 src geometry (normal gate)" if ngate else "", " + %g photometric" % extra_bytes_px if extra_bytes_px else "")},
                    "pose_err_vs_gt": {"t_m_max": e_t, "r_rad_max": e_r, "pairs_failed": int(nf)}}
        del d_fr, pin

    # c3: 1280x720, 256 independent pairs, ||t|| <= 3 cm, angle <= 2 deg
    pair_batch("c3_1280x720_256pairs", 1280, 720, 256, default_params(num_levels=3, iters=list(ITERS)),
               dict(seed=3), 0.0, "BASELINE configs[2]: 1280x720 synthetic depth, 256 independent frame pairs partitioned over the ranks")
    # c5: stress — 30 % invalid depth, motion up to 10 cm / 8 deg, Huber weights, more iterations
    noise = synth.Noise(p_invalid_pixel=0.21, p_invalid_block=0.12)
    P5 = default_params(num_levels=3, iters=[10, 10, 12], robust_kind=N.RST_ROBUST_HUBER, robust_scale=0.01, dist_max=0.6)
    pair_batch("c5_stress_1280x720_128pairs", 1280, 720, 128, P5, dict(seed=9, max_t=0.10, max_r=float(np.deg2rad(8.0)), noise=noise), 0.0,
               "BASELINE configs[4]: 30 % invalid depth (pixels + 16x16 blocks), motion up to 10 cm / 8 deg, Huber-weighted, iterations (10,10,12), "
               "128 pairs partitioned over the ranks")

    # c4: 848x480 RGB-D, frame-to-keyframe (keyframe every 10 frames), geometric + photometric (lambda = 0.5); weak scaling
    w, h = 848, 480
    intr = synth.intrinsics_for(w, h)
    n_frames = 61
    sc = synth.Scene(40 + rank)
    Twc = synth.trajectory(n_frames, seed=40 + rank, step_t=0.006, step_r=0.005)
    pin_d = torch.empty((n_frames, h, w), dtype=torch.int16, pin_memory=True)
    pin_c = torch.empty((n_frames, h, w, 3), dtype=torch.uint8, pin_memory=True)
    dep = pin_d.numpy().view(np.uint16); col = pin_c.numpy()
    for k in range(n_frames):
        _, c_k = sc.render(Twc[k], w, h, intr=intr, rgb=True, out=dep[k])
        col[k] = c_k
    src_slots = np.array([k for k in range(n_frames) if k % 10 != 0], dtype=np.int32)
    dst_slots = (src_slots // 10 * 10).astype(np.int32)
    n = len(src_slots)
    P4 = default_params(num_levels=3, iters=list(ITERS), photo_weight=0.5)
    al = Aligner(w, h, n_frames, n, device=local_rank, stream=stream.cuda_stream)
    al.begin(w, h, intr, P4)
    al.upload(dep, rgb=col)          # RGB-D frames come from the host; the upload is outside the timed region
    al.sync()
    d_poses = torch.empty((n, 16), dtype=torch.float32, device=dev)
    d_allp = torch.empty((world * n, 16), dtype=torch.float32, device=dev) if world > 1 else None

    def step4():
        al.preprocess(0, n_frames)
        al.align_slots(src_slots, dst_slots, fetch=False)
        al.copy_results_device(d_poses.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(d_allp, d_poses)
    steps = max(3, min(args.steps, 10))
    ms = time_steps(step4, steps)
    al.set_stream_split(0)
    al.profile_enable(True)
    time_steps(step4, 3)
    prof = al.profile_read()
    al.profile_enable(False)
    T4, st4 = al.align_slots(src_slots, dst_slots)
    errs = np.array([synth.pose_error(T4[i], synth.relative_pose(Twc[dst_slots[i]], Twc[src_slots[i]])) for i in range(n)])
    count = float(np.mean([s.count for s in st4]))
    t_l0 = prof.ms_icp[0] * 1e-3 / max(prof.launches_icp[0], 1)
    ach = n * (2.0 * w * h + (16.0 + 8.0) * count) / t_l0 / 1e9
    e_t, e_r, nf, frac = reduce_max([errs[:, 0].max(), errs[:, 1].max(), sum(1 for s in st4 if s.status != 0), ach / peak])
    al.close()
    out["c4_848x480_rgbd_f2kf"] = {
        "what": "BASELINE configs[3]: 848x480 RGB-D, every frame against its keyframe (one per 10 frames), geometric + photometric residual "
                "(lambda 0.5); %d frames / %d pairs per GPU, frames uploaded once outside the timed region (pre-processing + iterations timed)" % (n_frames, n),
        "width": w, "height": h, "pairs_per_gpu": n, "scaling": "weak", "value": world * n * steps / (ms * 1e-3), "unit": "pairs/s",
        "ms_per_step": ms / steps,
        "level0": {"launch_us": t_l0 * 1e6, "achieved_gbs": ach, "frac_of_measured_peak_max_over_ranks": frac, "associated_px_per_pair": count,
                   "bytes_per_px": "2 src depth + 16 gather + 8 photometric (src + dst intensity; SURVEY.md 8d)"},
        "pose_err_vs_gt": {"t_m_max": e_t, "r_rad_max": e_r, "pairs_failed": int(nf)}}
    return out


RESULT_OUT = sys.stdout


def main():
    # stdout carries exactly ONE JSON line: everything libraries print (NCCL's version banner, warnings)
    # is sent to stderr by pointing fd 1 at fd 2 and keeping a private handle on the real stdout
    global RESULT_OUT
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="headline regions only: no sustained / latency / H2D / extra-config legs (kernel experiments)")
    ap.add_argument("--no-split", action="store_true", help="single-stream schedule for the headline region too (profiling runs: launches of one kernel back to back)")
    ap.add_argument("--chunk", type=int, default=None, help="frames per upload/compute chunk of the e2e path")
    ap.add_argument("--size", default=None, help="WxH override for ad-hoc runs (e.g. 1280x720); the default is the BASELINE metric's 640x480")
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU per step override (default 129)")
    args = ap.parse_args()
    global W, H, FRAMES, METRIC
    if args.size:
        W, H = (int(x) for x in args.size.lower().split("x"))
        METRIC = f"icp_frame_pairs_per_sec_{W}x{H}"
    if args.frames:
        FRAMES = args.frames
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
