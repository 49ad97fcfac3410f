"""CPU, world_size 2 over gloo: the pair sharding + result gather used on the multi-GPU path.
The per-rank compute is injected (Oracle-N stands in for the rank's CUDA context here), so this
checks the host logic only: partitioning, ragged all-gather, bit-identical gathered results."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from realsensetracker_b200 import shard


def test_partition_is_a_contiguous_cover():
    for n in (0, 1, 5, 128, 257):
        for world in (1, 2, 3, 8):
            spans = [shard.partition(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard.sequence_partition(129, 2, 0) == (0, 65) and shard.sequence_partition(129, 2, 1) == (64, 129)


def _worker(rank, world, port, n_pairs, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        g = np.load(ROOT / "tests" / "golden" / "oracle_n_160x120.npz")
        f, intr = g["frames"], tuple(g["intr"])
        src = np.stack([f[1], f[2], f[2], f[0], f[1]][:n_pairs])
        dst = np.stack([f[0], f[1], f[0], f[1], f[2]][:n_pairs])
        P = O.default_params(iters=[2, 2, 1])

        def align_fn(s, d):
            res = [O.align_pair(s[i], d[i], intr, P) for i in range(len(s))]
            return np.stack([r[0] for r in res]) if res else np.zeros((0, 4, 4)), [r[1].status for r in res]

        poses, status = shard.align_pairs_sharded(align_fn, src, dst, world, rank)
        full, full_status = align_fn(src, dst) if rank == 0 else (None, None)
        q.put((rank, poses, status, full, full_status))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [4, 5])   # even split and ragged split
def test_sharded_result_equals_single_rank_bit_for_bit(n_pairs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_pairs
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, p0, s0, full, full_status), (_, p1, s1, _, _) = res
    assert np.array_equal(p0, p1) and np.array_equal(s0, s1)          # every rank holds the same gathered result
    assert np.array_equal(p0, np.asarray(full, dtype=np.float32).astype(np.float64))
    assert s0.tolist() == list(full_status)
