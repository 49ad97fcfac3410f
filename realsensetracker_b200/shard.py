"""Sharding of independent frame pairs across ranks (one process per GPU).

AlignIcp3d is a pure function of (src, dst, initial pose) with no shared state
(align_icp.cpp:73-161) and the reference processes pairs strictly one after another in one
thread (rs_replay_app.cpp:211-412), so pairs shard with NO collective on the data path. The only
communication is one all-gather of the per-pair results (64 B pose + 248 B statistics per pair).
Works with any torch.distributed backend: NCCL over NVLink on the B200 box, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def partition(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition: rank r gets [lo, hi). Sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sequence_partition(n_frames: int, world: int, rank: int) -> tuple[int, int]:
    """Frame range [lo, hi) of a sequence for rank r: the pair ranges are a block partition of the
    n_frames-1 frame-to-frame pairs, and consecutive ranks overlap by one frame so that each rank
    uploads and pre-processes each of its frames once."""
    plo, phi = partition(n_frames - 1, world, rank)
    return plo, phi + 1 if phi > plo else plo


def all_gather_rows(local, n_total: int, world: int, rank: int, group=None):
    """All-gathers row blocks of a 2-D tensor partitioned by `partition` (ragged allowed).
    Returns the [n_total, cols] tensor on every rank, rows in global pair order."""
    import torch
    import torch.distributed as dist
    sizes = [partition(n_total, world, r)[1] - partition(n_total, world, r)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank]
    if world == 1:
        return local
    if len(set(sizes)) == 1:
        out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def align_pairs_sharded(align_fn, src: np.ndarray, dst: np.ndarray, world: int, rank: int, device="cpu", group=None):
    """Runs `align_fn(src_block, dst_block) -> (poses [m,4,4], status [m])` on this rank's block of
    pairs and all-gathers poses and status words. `align_fn` is the rank's alignment context
    (Aligner.align_pairs on a GPU rank). Returns (poses [n,4,4] float64, status [n] int32)."""
    import torch
    n = src.shape[0]
    lo, hi = partition(n, world, rank)
    poses, status = align_fn(src[lo:hi], dst[lo:hi])
    # poses travel as their fp32 bit patterns so the gathered result is bit-identical to the local one
    loc = torch.from_numpy(np.ascontiguousarray(np.asarray(poses, dtype=np.float32).reshape(hi - lo, 16)).view(np.int32).copy())
    st = torch.from_numpy(np.asarray(status, dtype=np.int32).reshape(hi - lo, 1).copy())
    packed = torch.cat([loc, st], dim=1).to(device)
    allp = all_gather_rows(packed, n, world, rank, group).cpu().numpy()
    out = allp[:, :16].copy().view(np.float32).reshape(n, 4, 4).astype(np.float64)
    return out, allp[:, 16].copy()
