"""CPU, world_size 2 over gloo: the host-side sharding of independent pairs used by bench.py --gpus N.

There is no data-path collective: ranks take contiguous blocks of sequences, results are
gathered only for reporting.  This test runs the sharder with the oracle standing in for the device so the
N>1 plumbing (partition, max-over-ranks timing reduction, gather) is covered without a GPU.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cuahn_vio_b200.sharding import shard_range, reduce_max_ms
    n_seq, seq_len = 5, 3
    lo, hi = shard_range(n_seq, rank, world)
    mine = torch.arange(lo * seq_len, hi * seq_len, dtype=torch.float64)   # pair ids this rank would process
    ms = reduce_max_ms(10.0 + rank)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()]))
    q.put((rank, lo, hi, ms, [int(s) for s in sizes]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_everything():
    sys.path.insert(0, ROOT)
    from cuahn_vio_b200.sharding import shard_range
    for n in (1, 5, 8, 8192, 17):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, ms0, s0), (r1, lo1, hi1, ms1, s1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 3, 3, 5)          # contiguous blocks of sequences
    assert ms0 == ms1 == 11.0                             # max over ranks
    assert s0 == s1 == [9, 6]
