#!/usr/bin/env python
"""The only real collectives of the path, measured on hardware under torchrun (NCCL over NVLink): the policy-gradient
all-reduce (agents/graph_tsp_agent.py `_allreduce_gradients`, one flat f32 bucket) and the baseline t-test statistics
(`vrpx.sharding.paired_ttest_allreduce`, three doubles; reference graph_tsp_agent.py:299-306), plus a 2-epoch `train()`
on instance shards that checks every rank ends with identical weights and the same baseline decision.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_collectives.py
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vrp-gym_b200"), ROOT]
import numpy as np
import torch
import torch.distributed as dist

from agents import TSPAgent
from gym_vrp.envs import TSPEnv
from vrpx import sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
out = {"world": world}

# ---- gradient bucket: 1,154,432 f32 (TSP model), mean over ranks
n = 1154432
g = torch.full((n,), float(rank + 1), device=dev)
for _ in range(5):
    sharding.allreduce_mean_(g)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = ev(), ev()
e0.record()
iters = 50
for _ in range(iters):
    sharding.allreduce_mean_(g)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
g2 = torch.full((n,), float(rank + 1), device=dev)
sharding.allreduce_mean_(g2)
assert torch.allclose(g2, torch.full_like(g2, (world + 1) / 2.0)), "gradient mean over ranks is wrong"
out["grad_allreduce_ms"] = float(t.item())
out["grad_bucket_bytes"] = n * 4
out["grad_allreduce_busbw_gbs"] = 2.0 * (world - 1) / world * n * 4 / (float(t.item()) * 1e-3) / 1e9

# ---- baseline t-test sufficient statistics: same (mean, p) on every rank, equal to scipy on the concatenated samples
total = 65536 * world
rs = np.random.RandomState(0)
cm_all, cb_all = rs.rand(total) + 0.002, rs.rand(total)
b, e = sharding.shard_range(total, rank, world)
cm, cb = torch.tensor(cm_all[b:e], device=dev), torch.tensor(cb_all[b:e], device=dev)
for _ in range(3):
    mean, p = sharding.paired_ttest_allreduce(cm, cb)
torch.cuda.synchronize()
dist.barrier()
import time

w0 = time.perf_counter()
for _ in range(20):
    mean, p = sharding.paired_ttest_allreduce(cm, cb)
w = (time.perf_counter() - w0) / 20
from scipy import stats

t_ref, p_ref = stats.ttest_rel(cm_all, cb_all)
assert abs(mean - float((cm_all - cb_all).mean())) < 1e-12 and abs(p - p_ref) <= 1e-9 * max(p_ref, 1e-300) + 1e-15, (mean, p, p_ref)
tt = torch.tensor([w * 1e3], device=dev, dtype=torch.float64)
dist.all_reduce(tt, op=dist.ReduceOp.MAX)
out["ttest_stats_allreduce_ms_incl_host_sync"] = float(tt.item())
out["ttest_p"] = p

# ---- two REINFORCE epochs on instance shards: identical weights on every rank afterwards
env = TSPEnv(20, 512, 0, seed=5, instance_rng="philox", instance_offset=rank * 512)
agent = TSPAgent(seed=5, csv_path=os.path.join(tempfile.mkdtemp(), "log.csv"))
agent.train(env, epochs=2, eval_epochs=1, check_point_dir=tempfile.mkdtemp() + "/")
flat = torch.cat([p_.detach().reshape(-1) for p_ in agent.model.parameters()])
lo, hi = flat.clone(), flat.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
out["weights_identical_across_ranks"] = bool(torch.equal(lo, hi))
assert out["weights_identical_across_ranks"], "ranks diverged: the gradient all-reduce is not applied identically"
tflat = torch.cat([p_.detach().reshape(-1) for p_ in agent.target_model.parameters()])
lo, hi = tflat.clone(), tflat.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
out["baseline_identical_across_ranks"] = bool(torch.equal(lo, hi))
assert out["baseline_identical_across_ranks"]
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
