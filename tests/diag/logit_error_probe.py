"""Worst gap between the CUDA logits and the oracle at the chosen action, classic vs table mode (diagnostic, GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "vrp-gym_b200"), ROOT]
import numpy as np, torch
from agents import VRPAgent, TSPAgent
from gym_vrp.envs import VRPEnv, TSPEnv
from oracle import policy_oracle as po
from oracle.env_oracle import EnvOracle
N, B, seed = 20, 64, 1234
for tables in (False, True):
    env = VRPEnv(N, B, 1, seed)
    s = env.sampler
    xy, depot, demand = s.get_graph_positions(), s.get_depots()[:, 0], s.get_demands()[:, :, 0]
    agent = VRPAgent(seed=seed)
    agent.model.eval()
    agent.model.decoder.score_tables = tables
    with torch.no_grad():
        loss, _ = agent.model(env, rollout=True, want_logits=True)
    out = agent.model.last_rollout
    tape = out["tape"].cpu().numpy().astype(np.int64)
    got = out["logits"].cpu().numpy()
    sd = {k: v.float().cpu() for k, v in agent.model.state_dict().items()}
    loss_o, _, tr = po.rollout(sd, EnvOracle("vrp", xy, depot, demand), greedy=True, tape=tape, return_trace=True)
    ref = tr["logits"]
    fin = np.isfinite(ref)
    d = np.abs(got[fin] - ref[fin])
    print("tables", tables, "max abs logit diff", d.max(), "mean", d.mean(), "p99.9", np.quantile(d, 0.999))
    worst = 0
    for t in range(tape.shape[0]):
        for b in range(B):
            row = ref[t, b]
            gap = row[np.isfinite(row)].max() - row[tape[t, b]]
            worst = max(worst, gap)
    print("  worst oracle gap at chosen action", worst)
