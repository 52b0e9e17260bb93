"""Instance-sharded data parallelism for the rollout path: one process per GPU, every rank owns a
contiguous range of instances and treats it as its own reference batch, so the rollout needs NO collective.
Collectives (NCCL on GPUs, gloo in the CPU tests) appear only where the algorithm has a real exchange:

  * `allreduce_mean_` — the flat policy-gradient bucket, once per training step;
  * `paired_ttest_allreduce` — sufficient statistics (n, sum d, sum d^2) of the baseline t-test
    (agents/graph_tsp_agent.py:299-306), then the swap decision is identical on every rank;
  * `max_over_ranks` — device-timed durations for benchmarks (max over ranks, never wall clock).
"""
from __future__ import annotations

import math
from typing import Tuple

import torch


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `total` instances owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    """In-place mean over ranks of one flat bucket (gradients)."""
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(flat, op=d.ReduceOp.SUM)
        flat.div_(d.get_world_size())
    return flat


def max_over_ranks(values: torch.Tensor) -> torch.Tensor:
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(values, op=d.ReduceOp.MAX)
    return values


def paired_ttest_allreduce(cost_model: torch.Tensor, cost_baseline: torch.Tensor):
    """Two-sided paired t-test over ALL ranks' instances from 3 all-reduced doubles.
    Returns (mean difference, p-value); equals scipy.stats.ttest_rel on the concatenated samples."""
    diff = (cost_model - cost_baseline).double()
    stats = torch.stack([torch.tensor(float(diff.numel()), dtype=torch.float64, device=diff.device),
                         diff.sum(), (diff * diff).sum()])
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(stats, op=d.ReduceOp.SUM)
    n, s1, s2 = [float(x) for x in stats.tolist()]
    mean = s1 / n
    var = max((s2 - n * mean * mean) / (n - 1), 0.0) if n > 1 else 0.0
    if var == 0.0:
        return mean, (1.0 if mean == 0.0 else 0.0)
    t = mean / math.sqrt(var / n)
    from scipy import stats as sps

    return mean, float(2.0 * sps.t.sf(abs(t), n - 1))
