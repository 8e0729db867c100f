"""N > 1 host logic on CPU: world_size-2 gloo. Units (keyframes / sequences) are owned by exactly one rank,
there is no data-path collective, and the job number is (sum of units) / (max over ranks of the time)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from ngicp import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert sharding.world() == (rank, rank, world)
    bounds = np.cumsum([0] + [100 + 7 * i for i in range(9)])          # 9 keyframes of different sizes
    owned = sharding.units_for_rank(9, rank, world)
    off, slices = sharding.keyframe_segments(bounds, owned)
    local_points = off[-1]
    t_max, units = sharding.reduce_job(0.5 + rank, float(local_points))   # rank 1 is slower
    q.put((rank, owned, off, slices, t_max, units))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, own0, off0, sl0, t0, u0), (r1, own1, off1, sl1, t1, u1) = out
    assert sorted(own0 + own1) == list(range(9)) and not set(own0) & set(own1)      # every keyframe exactly once
    assert own0 == [0, 2, 4, 6, 8] and own1 == [1, 3, 5, 7]                         # keyframe i -> rank i mod G
    total = sum(100 + 7 * i for i in range(9))
    assert t0 == t1 == 1.5 and u0 == u1 == float(total)                             # max time, summed units on every rank
    assert off0[-1] + off1[-1] == total
    assert all(e - s == o2 - o1 for (s, e), o1, o2 in zip(sl0, off0[:-1], off0[1:]))


def test_single_process_needs_no_process_group():
    assert sharding.units_for_rank(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.reduce_job(2.0, 7.0) == (2.0, 7.0)
