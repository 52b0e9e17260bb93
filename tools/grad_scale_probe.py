"""How much precision do the f16-split backward GEMMs lose on small gradients?  The REINFORCE gradient is linear in the
advantage: grad(s * adv) / s must equal grad(adv).  Prints the largest relative deviation per parameter group for
s = 1e-3, 1e-5, 1e-7 (per-row gradients at a 65,536 batch are ~1e-5 of those of the 8-instance fixture)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vrp-gym_b200"), ROOT]
from agents import TSPAgent
from gym_vrp.envs import TSPEnv

z = np.load(os.path.join(ROOT, "tests", "golden", "policy_tsp_large.npz"))
key, N, B, seed = "50_8_31", 50, 8, 31
tape = z[key + "/tf_tape"]


def grads(scale):
    agent = TSPAgent(seed=seed)
    model = agent.model
    model.train()
    env = TSPEnv(N, B, 1, seed)
    loss_m, logp = model(env, rollout=False, tape=tape)
    baseline = torch.tensor(z[key + "/greedy_loss"], device=loss_m.device)
    adv = (loss_m - baseline) * -1
    model.zero_grad()
    model.backward(adv / B * scale)
    return {n: p.grad.detach().double().cpu() / scale for n, p in model.named_parameters() if p.grad is not None}


import vrpx.backward as vb

if len(sys.argv) > 1:
    vb._GRAD_TARGET_ENV = int(sys.argv[1])
    print("gain target 2^%d" % vb._GRAD_TARGET_ENV)
# reference: the fp32 SIMT cross-check path of the encoder backward (no f16 split)
def grads_simt():
    agent = TSPAgent(seed=seed)
    model = agent.model
    model.train()
    model.encoder.gemm_path = 1
    env = TSPEnv(N, B, 1, seed)
    loss_m, logp = model(env, rollout=False, tape=tape)
    baseline = torch.tensor(z[key + "/greedy_loss"], device=loss_m.device)
    model.zero_grad()
    model.backward((loss_m - baseline) * -1 / B)
    return {n: p.grad.detach().double().cpu() for n, p in model.named_parameters() if p.grad is not None}


ref = grads_simt()
gmax = max(g.abs().max().item() for g in ref.values())
for s in (1.0, 1e-3, 1e-5, 1e-7):
    got = grads(s)
    worst = {"encoder": 0.0, "decoder": 0.0}
    for n, g in ref.items():
        e = (got[n] - g).abs().max().item() / max(g.abs().max().item(), 1e-3 * gmax)
        k = "encoder" if n.startswith("encoder") else "decoder"
        worst[k] = max(worst[k], e)
    print(f"scale {s:g}: worst relative deviation encoder {worst['encoder']:.2e}, decoder {worst['decoder']:.2e}", flush=True)
