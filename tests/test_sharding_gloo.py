"""Host-side multi-rank logic on CPU: world_size-2 gloo.  Each rank processes its
contiguous shard of the query batch (with the CPU oracle standing in for the device
kernels -- this test is about the sharding / gather plumbing) and the gathered result
must equal the single-rank result."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, ret):
    import torch.distributed as dist

    for sub in ("mind-fcl_b200", "oracle"):
        sys.path.insert(0, os.path.join(ROOT, sub))
    import oracle_py
    import scenes
    import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes, pairs, p1, p2 = scenes.config_c2(n, np.float32, seed=5)
    b, e = sharding.shard_range(n, rank, world)
    oracle = oracle_py.PortOracle()
    d, _, _, ok = oracle.distance_batch(shapes, pairs[b:e], p1[b:e], p2[b:e])
    full = sharding.gather_counts(ok.astype(np.uint32), n, rank, world)
    if rank == 0:
        ret.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    sys.path.insert(0, os.path.join(ROOT, "mind-fcl_b200"))
    import sharding

    for n in (0, 1, 7, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_matches_single_rank(port_oracle):
    import scenes

    n = 4001
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, ret)) for r in range(2)]
    for p in procs:
        p.start()
    full = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shapes, pairs, p1, p2 = scenes.config_c2(n, np.float32, seed=5)
    _, _, _, ok = port_oracle.distance_batch(shapes, pairs, p1, p2)
    assert np.array_equal(full, ok.astype(np.uint32))
