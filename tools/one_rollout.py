#!/usr/bin/env python
"""Two greedy rollouts (one warm-up, one profiled) of a bench configuration — the short workload ncu captures run on.

    python tools/one_rollout.py [kind] [nodes] [batch] [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vrp-gym_b200"), ROOT]
import torch

from agents import IRPAgent, TSPAgent, VRPAgent
from gym_vrp.envs import IRPEnv, TSPEnv, VRPEnv

kind = sys.argv[1] if len(sys.argv) > 1 else "tsp"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
B = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
Env = {"tsp": TSPEnv, "vrp": VRPEnv, "irp": IRPEnv}[kind]
Agent = {"tsp": TSPAgent, "vrp": VRPAgent, "irp": IRPAgent}[kind]
env = Env(N, B, 0, seed=69, instance_rng="philox")
agent = Agent(seed=69)
for _ in range(reps):
    env.restart_episode()
    loss = agent.evaluate(env)
torch.cuda.synchronize()
print("mean cost", -loss.mean().item(), "steps", env.step_count)
