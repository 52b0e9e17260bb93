#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (all small, committed):
  random_agent_costs.npz  the 6,912 Random-Agent costs of reference reproduction_log/*.csv
  env_tapes.npz           per-step transitions of reference TSPEnv/VRPEnv/IRPEnv under recorded actions
  policy_<kind>.npz       embeddings / per-step logits / greedy tapes / costs / teacher-forced log-probs
                          from reference {TSP,VRP,IRP}Agent(seed) with seed-initialised weights
  policy_<kind>_large.npz the same at the benchmarked node counts N = 40 / 50 / 100 (B = 8)
  known_answers.json      constants asserted by the reference's own tests + SURVEY App. C
"""
import csv
import json
import os
import sys
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, REF)

import torch  # noqa: E402

import agents.graph_decoder as ref_decoder_mod  # noqa: E402
from agents import IRPAgent, RandomAgent, TSPAgent, VRPAgent  # noqa: E402
from gym_vrp.envs import IRPEnv, TSPEnv, VRPEnv  # noqa: E402

ENVS = {"tsp": TSPEnv, "vrp": VRPEnv, "irp": IRPEnv}
AGENTS = {"tsp": TSPAgent, "vrp": VRPAgent, "irp": IRPAgent}


# ---------------------------------------------------------------- 1. CSV golden costs
def golden_random_costs():
    out = {}
    for n in (20, 30, 40):
        for kind in ("TSP", "VRP", "IRP"):
            path = f"{REF}/reproduction_log/reproduction_results_{n}_nodes_model_{kind}.csv"
            per_seed = {}
            with open(path) as f:
                for row in csv.DictReader(f):
                    if row["Model"].endswith("Random-Agent"):
                        per_seed.setdefault(int(row["Seed"]), []).append(float(row["Mean Distance"]))
            for seed, vals in per_seed.items():
                assert len(vals) == 256, (path, seed, len(vals))
                out[f"{kind.lower()}_{n}_{seed}"] = np.asarray(vals, dtype=np.float64)
    assert sum(v.size for v in out.values()) == 6912
    np.savez_compressed(os.path.join(HERE, "random_agent_costs.npz"), **out)
    # self-check: the reference reproduces them here (float32-exact)
    env = TSPEnv(20, 256, 3, seed=1234)
    r = RandomAgent(1234)(deepcopy(env)).numpy()
    assert np.array_equal(np.float32(out["tsp_20_1234"]), -r), "reference no longer reproduces its CSV"
    print("random_agent_costs.npz:", len(out), "series")


# ---------------------------------------------------------------- 2. env transition tapes
def env_tapes():
    out = {}
    rs = np.random.RandomState(7)
    for kind in ("tsp", "vrp", "irp"):
        for (N, B, seed) in ((5, 8, 11), (13, 16, 22), (20, 32, 1234), (33, 8, 5)):
            env = ENVS[kind](N, B, min(3, B), seed)
            key = f"{kind}_{N}_{B}_{seed}"
            xy = env.sampler.get_graph_positions()
            out[key + "/xy"] = xy
            out[key + "/depot"] = env.depots[:, 0].astype(np.int64)
            out[key + "/demand"] = env.sampler.get_demands()[:, :, 0]
            out[key + "/draw_idxs"] = env.draw_idxs
            st = env.get_state()
            load = None
            if kind == "irp":
                st, load = st
            out[key + "/state0"] = st.copy()
            acts, vis, masks, loads, rews, dones = [], [], [], [], [], []
            done = False
            while not done:
                mask = st[:, :, -1]
                a = np.array([rs.choice(np.flatnonzero(mask[b] == 0)) for b in range(B)])
                st, r, done, _ = env.step(a[:, None])
                if kind == "irp":
                    st, load = st
                    loads.append(load.copy())
                acts.append(a)
                vis.append(env.visited.copy())
                masks.append(st[:, :, -1].copy())
                rews.append(r.copy())
                dones.append(bool(done))
            out[key + "/actions"] = np.stack(acts).astype(np.int64)
            out[key + "/visited"] = np.stack(vis).astype(np.uint8)
            out[key + "/mask"] = np.stack(masks).astype(np.uint8)
            out[key + "/reward"] = np.stack(rews)
            out[key + "/done"] = np.asarray(dones)
            if kind == "irp":
                out[key + "/load"] = np.stack(loads)
            # second episode after reset(): the stream continues without reseeding (tsp.py:150-160)
            st = env.reset()
            if kind == "irp":
                st = st[0]
            out[key + "/reset_xy"] = env.sampler.get_graph_positions()
            out[key + "/reset_depot"] = env.depots[:, 0].astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "env_tapes.npz"), **out)
    print("env_tapes.npz:", len(out), "arrays")


# ---------------------------------------------------------------- 3. policy traces
class _Recorder:
    """Captures the masked pointer logits `u` (graph_decoder.py:98) by wrapping masked_fill."""

    def __init__(self):
        self.logits = []
        self._orig = torch.Tensor.masked_fill

    def __enter__(self):
        rec = self

        def wrapped(t, mask, value):
            res = rec._orig(t, mask, value)
            if isinstance(value, float) and value == float("-inf") and t.dim() == 3 and t.shape[1] == 1:
                rec.logits.append(res.detach()[:, 0].clone().numpy())
            return res

        torch.Tensor.masked_fill = wrapped
        return self

    def __exit__(self, *a):
        torch.Tensor.masked_fill = self._orig


class _ReplayCategorical(ref_decoder_mod.Categorical):
    tape = None
    t = 0

    def sample(self, *a, **k):
        cls = _ReplayCategorical
        act = torch.as_tensor(cls.tape[cls.t], dtype=torch.long)[:, None]
        cls.t += 1
        return act


def _weights_checksum(sd):
    return float(sum(v.double().sum().item() for v in sd.values()))


SMALL_CASES = ((4, 2, 69), (10, 8, 7), (20, 32, 1234))
# the node counts the benchmarks run (BASELINE.json configs 3-5: N = 40 / 50 / 100) at a batch the reference plays in seconds
LARGE_CASES = ((40, 8, 30), (50, 8, 31), (100, 8, 32))


def policy_traces(cases=SMALL_CASES, suffix=""):
    for kind in ("tsp", "vrp", "irp"):
        out = {}
        for (N, B, seed) in cases:
            key = f"{N}_{B}_{seed}"
            env = ENVS[kind](N, B, 1, seed)
            agent = AGENTS[kind](seed=seed)
            out[key + "/wsum"] = np.float64(_weights_checksum(agent.model.state_dict()))
            embs = []
            hook = agent.model.encoder.register_forward_hook(lambda m, i, o: embs.append(o.detach().numpy().copy()))
            # greedy eval on the constructor instances (reproduction.py:47 path)
            acts = []
            env_g = deepcopy(env)

            def rec_step(a, _e=env_g, _acts=acts):
                _acts.append(np.asarray(a)[:, 0].copy())
                return type(_e).step(_e, a)

            env_g.step = rec_step  # instance attribute shadows the method
            with _Recorder() as rec:
                cost = agent.evaluate(env_g)
            out[key + "/greedy_actions"] = np.stack(acts).astype(np.int64)
            out[key + "/greedy_logits"] = np.stack(rec.logits).astype(np.float32)
            out[key + "/greedy_loss"] = cost.numpy()
            out[key + "/emb_eval"] = embs[-1][: min(B, 4)] if not suffix else embs[-1]
            # teacher-forced sampled-mode log-probs (eval-mode BN) along a random feasible tape
            env_t = deepcopy(env)
            rs = np.random.RandomState(3)
            tape = []
            e2 = deepcopy(env)
            st = e2.get_state()
            done = False
            while not done:
                if kind == "irp":
                    st = st[0]
                m = st[:, :, -1]
                a = np.array([rs.choice(np.flatnonzero(m[b] == 0)) for b in range(B)])
                tape.append(a)
                st, _, done, _ = e2.step(a[:, None])
            tape = np.stack(tape)
            _ReplayCategorical.tape, _ReplayCategorical.t = tape, 0
            ref_decoder_mod.Categorical = _ReplayCategorical
            try:
                agent.model.eval()
                with torch.no_grad(), _Recorder() as rec2:
                    loss_s, logp_s = agent.model(env_t, rollout=False)
            finally:
                ref_decoder_mod.Categorical = _ReplayCategorical.__mro__[1]
            out[key + "/tf_tape"] = tape.astype(np.int64)
            out[key + "/tf_logits"] = np.stack(rec2.logits).astype(np.float32)
            out[key + "/tf_loss"] = loss_s.numpy()
            out[key + "/tf_logp"] = logp_s.numpy()
            # train-mode (batch-statistics BN) embeddings + teacher-forced log-probs + REINFORCE grads
            env_tr = deepcopy(env)
            agent2 = AGENTS[kind](seed=seed)
            embs2 = []
            agent2.model.encoder.register_forward_hook(lambda m, i, o: embs2.append(o.detach().numpy().copy()))
            _ReplayCategorical.tape, _ReplayCategorical.t = tape, 0
            ref_decoder_mod.Categorical = _ReplayCategorical
            try:
                agent2.model.train()
                loss_m, logp_m = agent2.model(env_tr, rollout=False)
            finally:
                ref_decoder_mod.Categorical = _ReplayCategorical.__mro__[1]
            baseline = torch.tensor(out[key + "/greedy_loss"])
            adv = (loss_m - baseline) * -1  # graph_tsp_agent.py:179
            loss = (adv * logp_m).mean()  # :180
            agent2.opt.zero_grad()
            loss.backward()
            out[key + "/train_emb"] = embs2[-1][: min(B, 4)] if not suffix else embs2[-1]
            out[key + "/train_logp"] = logp_m.detach().numpy()
            out[key + "/train_loss"] = np.float32(loss.item())
            for name, p in agent2.model.named_parameters():
                if p.grad is None:
                    continue
                g = p.grad.detach().reshape(-1)
                out[key + "/grad_norm/" + name] = np.float64(g.double().norm().item())
                out[key + "/grad_head/" + name] = g[:16].numpy().copy()
            bn = agent2.model.encoder.attention_layers[0].bn1.norm
            out[key + "/bn_l0_running_mean"] = bn.running_mean.numpy().copy()
            out[key + "/bn_l0_running_var"] = bn.running_var.numpy().copy()
            hook.remove()
        np.savez_compressed(os.path.join(HERE, f"policy_{kind}{suffix}.npz"), **out)
        print(f"policy_{kind}{suffix}.npz:", len(out), "arrays")


# ---------------------------------------------------------------- 4. known answers
def known_answers():
    ka = {
        # reference tests/test_agent.py:69,84,99,114 (session seed 69)
        "test_random_agent_mean": -5.585874557495117,
        "test_tsp_agent_mean": -1.5130789279937744,
        "test_vrp_agent_mean": -1.952601671218872,
        "test_irp_agent_mean": -2.9770922660827637,
    }
    # SURVEY App. C, recomputed here from the reference
    np.random.seed(69)
    e = VRPEnv(3, 2, 2)
    ka["seed69_vrp_3_2_depots"] = e.depots[:, 0].tolist()
    ka["seed69_vrp_3_2_draw_idxs"] = e.draw_idxs.tolist()
    for kind in ("tsp", "vrp", "irp"):
        env = ENVS[kind](20, 256, 3, seed=1234)
        ag = AGENTS[kind](seed=1234)
        loss = ag.evaluate(env)
        ka[f"{kind}_20_256_1234_greedy_mean"] = float(loss.mean().item())
        ka[f"{kind}_20_256_1234_steps"] = int(env.step_count)
        ka[f"{kind}_20_256_1234_wsum"] = _weights_checksum(ag.model.state_dict())
    env = TSPEnv(20, 256, 3, seed=1234)
    ka["tsp_20_256_1234_draw_idxs"] = env.draw_idxs.tolist()
    ka["tsp_20_256_1234_depots8"] = env.depots[:8, 0].tolist()
    ka["tsp_20_256_1234_xy00"] = env.sampler.get_graph_positions()[0, 0].tolist()
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(ka, f, indent=1)
    print("known_answers.json:", ka)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "large":
        policy_traces(LARGE_CASES, "_large")
        sys.exit(0)
    golden_random_costs()
    env_tapes()
    policy_traces()
    policy_traces(LARGE_CASES, "_large")
    known_answers()
