#!/usr/bin/env python
"""Train small checkpoints WITH THE UNMODIFIED REFERENCE (CPU) and record its greedy evaluation on fixed instances.

The reference's published `check_points/**/model_epoch_850.pt` blobs are not in the tree (`.MISSING_LARGE_BLOBS`), so
trained-checkpoint parity (north_star: mean tour cost within 0.1 %) is pinned on checkpoints produced here by the
reference's own `Agent.train` (agents/graph_tsp_agent.py:150-208), the substitute SURVEY §8c-(ii) names.  Build container
only (needs /root/reference):

    python tests/golden/make_checkpoints.py tsp 20 123 250      # kind, nodes, seed, epochs

Outputs (committed): tests/golden/ckpt_<kind>_<N>_<seed>.pt  (reference `model.state_dict()`, reproduction.py:41-42 format)
                     tests/golden/ckpt_<kind>_<N>_<seed>_eval.npz  (greedy tapes / logits / costs of the reference with that
                     checkpoint on Env(N, 256, 3, seed=1234) and on Env(2N, 64, 3, seed=2468), plus the training log)
"""
import os
import sys
import tempfile
import time
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, REF)

import torch  # noqa: E402

from agents import IRPAgent, TSPAgent, VRPAgent  # noqa: E402
from gym_vrp.envs import IRPEnv, TSPEnv, VRPEnv  # noqa: E402

sys.path.insert(0, HERE)
from make_golden import _Recorder  # noqa: E402

ENVS = {"tsp": TSPEnv, "vrp": VRPEnv, "irp": IRPEnv}
AGENTS = {"tsp": TSPAgent, "vrp": VRPAgent, "irp": IRPAgent}


def record_eval(agent, env):
    """agent.evaluate(env) (reproduction.py:47) with the per-step actions and masked pointer logits recorded."""
    acts = []

    def rec_step(a, _e=env, _acts=acts):
        _acts.append(np.asarray(a)[:, 0].copy())
        return type(_e).step(_e, a)

    env.step = rec_step
    with _Recorder() as rec:
        loss = agent.evaluate(env)
    return np.stack(acts).astype(np.int64), np.stack(rec.logits).astype(np.float32), loss.numpy()


def main():
    kind, N, seed, epochs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    torch.set_num_threads(int(os.environ.get("THREADS", "3")))
    tmp = tempfile.mkdtemp()
    env = ENVS[kind](num_nodes=N, batch_size=256, seed=seed)              # train_models.py:10-12
    agent = AGENTS[kind](seed=seed, csv_path=os.path.join(tmp, "loss.csv"))  # train_models.py:14-16
    t0 = time.time()
    agent.train(env, epochs=epochs, check_point_dir=os.path.join(tmp, "ck") + "/")
    print(f"trained {kind}-{N} seed {seed} for {epochs} epochs in {time.time() - t0:.0f} s", flush=True)
    stem = os.path.join(HERE, f"ckpt_{kind}_{N}_{seed}")
    torch.save(agent.model.state_dict(), stem + ".pt")
    log = np.loadtxt(os.path.join(tmp, "loss.csv"), delimiter=",", skiprows=1)

    out = {"train_log": log, "epochs": np.int64(epochs)}
    # reload through the reference's own loading path (reproduction.py:41-42) into a fresh agent
    agent2 = AGENTS[kind](seed=seed)
    agent2.model.load_state_dict(torch.load(stem + ".pt"))
    for tag, (n2, b2, s2) in {"a": (N, 256, 1234), "b": (2 * N, 64, 2468)}.items():
        e = ENVS[kind](n2, b2, 3, seed=s2)
        out[f"{tag}/cfg"] = np.asarray([n2, b2, s2])
        acts, logits, loss = record_eval(agent2, deepcopy(e))
        out[f"{tag}/greedy_actions"], out[f"{tag}/greedy_logits"], out[f"{tag}/greedy_loss"] = acts, logits, loss
        print(f"  eval {kind}-{n2} B={b2} seed {s2}: mean cost {-loss.mean():.6f}, steps {acts.shape[0]}", flush=True)
    np.savez_compressed(stem + "_eval.npz", **out)


if __name__ == "__main__":
    main()
