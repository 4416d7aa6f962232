"""N>1 host logic on CPU: world_size-2 `gloo` processes.

The path shards by contiguous point ranges with NO data-path collective:
every rank owns gstools_core.shard_bounds(M, world, rank) and would run the CUDA path on it.  There
is no GPU here, so each rank evaluates its shard with the CPU oracle (test infrastructure) as the
stand-in compute; the test checks what the multi-rank plumbing is responsible for:
  * shards are disjoint, contiguous, cover [0, M), and agree between the C ABI and every rank;
  * concatenating the per-rank results reproduces the single-process result bit for bit
    (mode order per point does not depend on the partition);
  * the max-over-ranks timing reduction bench.py uses.
"""
import os
import socket

import numpy as np
import pytest

import gstools_core as gc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, m, q):
    import torch
    import torch.distributed as dist

    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)                      # same inputs on every rank
    k = rng.normal(size=(3, 64)); z1 = rng.normal(size=64); z2 = rng.normal(size=64)
    pos = rng.uniform(-10, 10, size=(3, m))
    j0, j1 = gc.shard_bounds(m, world, rank)
    local = oracle.summate(k, z1, z2, np.ascontiguousarray(pos[:, j0:j1]))
    # bookkeeping exchange only (bounds + timing), never the field data path
    bounds = [None] * world
    dist.all_gather_object(bounds, (j0, j1))
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    q.put((rank, j0, j1, local, bounds, float(t.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("m", [5000, 4096, 70001])
def test_two_rank_sharding(m):
    import torch.multiprocessing as mp

    import oracle

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, a1, out0, b0, t0), (r1, c0, c1, out1, b1, t1) = res
    assert (a0, c1) == (0, m) and a1 == c0                        # contiguous cover of [0, M)
    assert a1 % 1024 == 0 or a1 == m                              # boundary on a chunk multiple
    assert b0 == b1 == [(a0, a1), (c0, c1)]                       # every rank sees the same partition
    assert t0 == t1 == 2.0                                        # max over ranks
    rng = np.random.default_rng(3)
    k = rng.normal(size=(3, 64)); z1 = rng.normal(size=64); z2 = rng.normal(size=64)
    pos = rng.uniform(-10, 10, size=(3, m))
    full = oracle.summate(k, z1, z2, pos)
    assert np.array_equal(np.concatenate([out0, out1]), full)


def test_shard_bounds_properties():
    for m in (0, 1, 1023, 1024, 10**6, 10**8 + 7):
        for g in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(g):
                b, e = gc.shard_bounds(m, g, r)
                assert b == prev and e >= b
                prev = e
            assert prev == m
    with pytest.raises(ValueError):
        gc.shard_bounds(10, 2, 2)
