"""CPU: the N>1 host logic (instance sharding, max-over-ranks timing) with world_size 2 over gloo."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from abc_b200.sharding import instance_range, max_over_ranks, sum_over_ranks


def test_instance_range_partitions_exactly():
    for n in (0, 1, 7, 256, 10000):
        for world in (1, 2, 3, 4, 8):
            got = [instance_range(n, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        instance_range(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = instance_range(10000, world, rank)
    ms = 10.0 + rank                      # pretend device time of this rank
    mx = max_over_ranks([ms, float(hi - lo)], dist)
    tot = sum_over_ranks([float(hi - lo)], dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, lo, hi, mx, tot))


def test_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert [(r[1], r[2]) for r in res] == [(0, 5000), (5000, 10000)]
    for r in res:
        assert r[3] == [11.0, 5000.0]     # max over ranks, identical on both
        assert r[4] == [10000.0]          # all instances accounted for
