"""GPU, world_size 2 over NCCL: pairs sharded over two B200s, one alignment context per rank, results gathered with
one all_gather — must equal the single-context result BIT FOR BIT (pairs are independent and every pair is reduced in
image-size-determined blocks, so neither the partition nor the rank changes a bit). Needs two devices
(`gpurun --gpus 2 -- python -m pytest tests/test_sharding_nccl.py -m gpu`); skipped on a one-GPU box, where the gloo
test (tests/test_sharding_gloo.py) covers the host logic."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, n_pairs, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from realsensetracker_b200 import Aligner, default_params, shard, synth
        w, h = 640, 480
        intr = synth.intrinsics_for(w, h)
        src, dst, gt = synth.render_pairs(n_pairs, w, h, seed=21)
        P = default_params()
        al = Aligner(w, h, 2 * n_pairs, n_pairs, device=rank)

        def align_fn(s, d):
            if len(s) == 0:
                return np.zeros((0, 4, 4)), []
            T, st = al.align_pairs(s, d, intr, P)
            return T, [x.status for x in st]

        poses, status = shard.align_pairs_sharded(align_fn, src, dst, world, rank, device=torch.device("cuda", rank))
        full, full_status = align_fn(src, dst) if rank == 0 else (None, None)
        err = max(synth.pose_error(poses[i], gt[i])[0] for i in range(n_pairs))
        al.close()
        q.put((rank, poses, status, full, full_status, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two CUDA devices (NCCL refuses two ranks on one device)")
@pytest.mark.parametrize("n_pairs", [6, 7])   # even and ragged split
def test_nccl_sharded_equals_single_context_bit_for_bit(n_pairs):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000) + n_pairs
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, p0, s0, full, full_status, err0), (_, p1, s1, _, _, _) = res
    assert np.array_equal(p0, p1) and np.array_equal(s0, s1)          # every rank holds the same gathered result
    assert np.array_equal(p0, np.asarray(full, dtype=np.float32).astype(np.float64))   # == one context doing all pairs
    assert s0.tolist() == list(full_status) and all(s == 0 for s in s0.tolist())
    assert err0 < 1.5e-3
