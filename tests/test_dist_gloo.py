"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic (vrpx/sharding.py): shard ranges, the gradient
bucket all-reduce, the baseline t-test from all-reduced sufficient statistics, max-over-ranks timing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "vrp-gym_b200"))
    import torch.distributed as dist
    from vrpx import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 1001
    b, e = sharding.shard_range(total, rank, world)
    rs = np.random.RandomState(0)
    cm = torch.tensor(rs.rand(total) + 0.02)
    cb = torch.tensor(rs.rand(total))
    mean, p = sharding.paired_ttest_allreduce(cm[b:e], cb[b:e])
    g = torch.full((10,), float(rank + 1))
    sharding.allreduce_mean_(g)
    t = sharding.max_over_ranks(torch.tensor([1.0 + rank, 5.0 - rank]))
    q.put((rank, b, e, mean, p, g.tolist(), t.tolist()))
    dist.destroy_process_group()


def test_sharding_and_collectives_gloo():
    from scipy import stats

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, b0, e0, m0, p0, g0, t0), (r1, b1, e1, m1, p1, g1, t1) = res
    assert (b0, e0, b1, e1) == (0, 501, 501, 1001)          # contiguous, sizes differ by at most one
    rs = np.random.RandomState(0)
    cm, cb = rs.rand(1001) + 0.02, rs.rand(1001)
    t_ref = stats.ttest_rel(cm.tolist(), cb.tolist())
    assert m0 == m1 and p0 == p1                              # same decision on every rank
    assert np.isclose(m0, (cm - cb).mean()) and np.isclose(p0, t_ref.pvalue, rtol=1e-9)
    assert g0 == g1 == [1.5] * 10
    assert t0 == t1 == [2.0, 5.0]


def test_shard_range_covers_everything():
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vrp-gym_b200"))
    from vrpx.sharding import shard_range

    for total, world in ((1 << 20, 8), (1001, 8), (7, 8), (65536, 3)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
