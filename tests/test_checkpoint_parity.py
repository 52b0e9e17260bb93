"""Trained-checkpoint parity (north_star: "mean tour cost matches within 0.1 % using the repo's check_points").

The published `check_points/**/model_epoch_850.pt` blobs are absent from the reference tree (.MISSING_LARGE_BLOBS), so
the checkpoints are produced by the UNMODIFIED reference's own `Agent.train` on CPU (tests/golden/make_checkpoints.py:
TSP-20 / VRP-20, seed 123, B = 256 as train_models.py:4-6; 300 / 250 epochs -> mean cost 4.38 / 4.61, random 9.9 / 11.6),
saved with `torch.save(model.state_dict())` exactly like graph_tsp_agent.py:210-225, and evaluated by the reference
(reproduction.py:41-47) on Env(20, 256, 3, seed=1234) and — the 20-in-40 generalisation run of reproduction.sh —
Env(40, 64, 3, seed=2468).  Here:
  * CPU: the oracle with the trained weights reproduces the reference's logits and tours (pins the oracle off
    seed-initialised weights); the reference loads a checkpoint THIS repo saved (container only);
  * GPU: the reference-saved state_dict loads unchanged, mean cost within 0.1 %, tours identical modulo near-ties,
    teacher-forced logits within 5 x the reference's own fp32 error against a float64 evaluation (`_check_trained_logits`)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPTS = [("tsp", 20, 123), ("vrp", 20, 123)]


def _cls(kind):
    from agents import IRPAgent, TSPAgent, VRPAgent
    from gym_vrp.envs import IRPEnv, TSPEnv, VRPEnv

    return {"tsp": (TSPEnv, TSPAgent), "vrp": (VRPEnv, VRPAgent), "irp": (IRPEnv, IRPAgent)}[kind]


def _paths(golden_dir, kind, N, seed):
    stem = os.path.join(golden_dir, f"ckpt_{kind}_{N}_{seed}")
    return stem + ".pt", stem + "_eval.npz"


def _rel(got, ref):
    return float((np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)).max())


def _f64_logits(kind, sd, xy, depot, demand, tape):
    """The same f32 weights on the same f32-rounded observations, evaluated in float64 (oracle): the yardstick."""
    from oracle import policy_oracle as po
    from oracle.env_oracle import EnvOracle

    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    _, _, tr = po.rollout(sd64, EnvOracle(kind, xy, depot, demand), greedy=True, tape=tape, return_trace=True,
                          dtype=torch.float64)
    return tr["logits"]


def _check_trained_logits(got, ref32, exact, factor, what):
    """Logit criterion on TRAINED weights.  The trained policy is sharp (large weights, peaked glimpse softmax, pointer
    dot products with heavy cancellation): the unmodified reference's own fp32 logits sit 1.5e-5 .. 1.9e-5 (relative,
    floor 1) away from the float64 value of the same network, so "within 1e-5 of the reference" is not defined there.
    Both implementations are therefore measured against the float64 yardstick: the error of `got` must stay within
    `factor` x the reference's own fp32 error on this very trace (seed-initialised weights keep the plain 1e-5
    everywhere else in the suite).  factor: 2 for the fp32 oracle; 5 for the CUDA path — measured 2.4 .. 3.7
    (profiles/r02_summary.md): its tensor-core contractions truncate when they accumulate (csrc/gemm_tc4.cu, SPLITACC),
    which the cancellation in the trained pointer logits amplifies."""
    fin = np.isfinite(ref32)
    assert np.array_equal(fin, np.isfinite(got)), what
    noise = _rel(ref32[fin].astype(np.float64), exact[fin])
    err = _rel(got[fin].astype(np.float64), exact[fin])
    assert noise < 2.5e-5, (what, noise)
    assert err <= max(1e-5, factor * noise), (what, err, noise, err / noise)
    assert _rel(got[fin], ref32[fin]) < 1.2e-4, what
    return err, noise


@pytest.mark.parametrize("kind,N,seed", CKPTS)
def test_oracle_with_trained_checkpoint(golden_dir, kind, N, seed):
    from oracle import policy_oracle as po
    from oracle.env_oracle import EnvOracle, seeded_env_instances

    pt, ev = _paths(golden_dir, kind, N, seed)
    sd = {k: v.float() for k, v in torch.load(pt, map_location="cpu").items()}
    z = np.load(ev)
    assert float(z["train_log"][-10:, 2].mean()) > -5.2, "the checkpoint is a trained one (cost well below random)"
    for tag in ("a", "b"):
        n2, b2, s2 = (int(v) for v in z[f"{tag}/cfg"])
        _, xy, depot, demand = seeded_env_instances(n2, b2, 3, s2)
        tape, ref = z[f"{tag}/greedy_actions"], z[f"{tag}/greedy_logits"]
        loss, _, tr = po.rollout(sd, EnvOracle(kind, xy, depot, demand), greedy=True, tape=tape, return_trace=True)
        _check_trained_logits(tr["logits"], ref, _f64_logits(kind, sd, xy, depot, demand, tape), 2.0, (kind, tag))
        assert np.allclose(loss.numpy(), z[f"{tag}/greedy_loss"], rtol=1e-6, atol=1e-6)


def test_state_dict_layout_equals_reference_checkpoint(golden_dir):
    """Keys, shapes and dtypes of this repo's models equal the reference-saved state_dict:
    `load_state_dict(strict=True)` in either direction."""
    for kind, N, seed in CKPTS:
        pt, _ = _paths(golden_dir, kind, N, seed)
        ref = torch.load(pt, map_location="cpu")
        own = _cls(kind)[1](seed=seed).model.state_dict()
        assert list(own.keys()) == list(ref.keys())
        for k in ref:
            assert own[k].shape == ref[k].shape and own[k].dtype == ref[k].dtype, k
        _cls(kind)[1](seed=seed).model.load_state_dict(ref, strict=True)


@pytest.mark.reference
@pytest.mark.parametrize("kind", ["tsp", "vrp", "irp"])
def test_reference_loads_checkpoint_saved_here(tmp_path, kind):
    """Reverse direction (graph_tsp_agent.py:210-225 -> reproduction.py:41-42): a checkpoint written by THIS repo's
    agent loads, strictly, into the unmodified reference agent and gives the same weights."""
    agent = _cls(kind)[1](seed=77)
    path = str(tmp_path / "ck") + "/"
    os.makedirs(path)
    # save_model writes every 50th epoch (not 0), like the reference
    agent.save_model(episode=50, check_point_dir=path)
    agent.save_model(episode=0, check_point_dir=path)
    agent.save_model(episode=7, check_point_dir=path)
    assert os.listdir(path) == ["model_epoch_50.pt"]
    wsum = float(sum(v.double().sum().item() for v in agent.model.state_dict().values()))
    code = (
        "import sys, torch\n"
        "from agents import TSPAgent, VRPAgent, IRPAgent\n"
        f"a = {{'tsp': TSPAgent, 'vrp': VRPAgent, 'irp': IRPAgent}}['{kind}'](seed=1)\n"
        f"a.model.load_state_dict(torch.load('{path}model_epoch_50.pt'))\n"
        "print('WSUM', repr(float(sum(v.double().sum().item() for v in a.model.state_dict().values()))))\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle", "stubs") + ":/root/reference")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = float(out.stdout.split("WSUM")[1].strip())
    assert got == wsum


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,seed", CKPTS)
def test_gpu_trained_checkpoint_costs_and_tours(golden_dir, kind, N, seed):
    Env, Agent = _cls(kind)
    pt, ev = _paths(golden_dir, kind, N, seed)
    z = np.load(ev)
    for tag in ("a", "b"):
        n2, b2, s2 = (int(v) for v in z[f"{tag}/cfg"])
        env = Env(n2, b2, 3, seed=s2)                                # reproduction.py:32-34
        agent = Agent(seed=s2)                                       # :40
        agent.model.load_state_dict(torch.load(pt, map_location=agent.device))   # :41-42, unchanged file
        loss = agent.evaluate(env)                                   # :47
        ref_loss, ref_tape, ref_logits = z[f"{tag}/greedy_loss"], z[f"{tag}/greedy_actions"], z[f"{tag}/greedy_logits"]
        got_mean, ref_mean = float(loss.mean().item()), float(ref_loss.mean())
        assert abs(got_mean - ref_mean) <= 1e-3 * abs(ref_mean), (kind, tag, got_mean, ref_mean)   # the 0.1 % criterion
        tape = agent.model.last_rollout["tape"].cpu().numpy()
        T = min(tape.shape[0], ref_tape.shape[0])
        same = 0
        for b in range(b2):
            diff = np.flatnonzero(tape[:T, b] != ref_tape[:T, b])
            if diff.size:   # documented near-tie at the first divergence
                t = diff[0]
                top2 = np.sort(ref_logits[t, b][np.isfinite(ref_logits[t, b])])[-2:]
                assert top2[1] - top2[0] < 2e-5, (kind, tag, b, t, top2)
            else:
                same += 1
                assert abs(loss[b].item() - ref_loss[b]) <= 1e-5 * max(1.0, abs(ref_loss[b]))
        assert same >= 0.98 * b2, (kind, tag, same)
        # teacher-forced along the reference's tape: every logit within 1e-5
        env2 = Env(n2, b2, 3, seed=s2)
        with torch.no_grad():
            agent.model(env2, rollout=True, tape=ref_tape, want_logits=True)
        got = agent.model.last_rollout["logits"].cpu().numpy()
        s = env2.sampler
        sd = {k: v.float().cpu() for k, v in agent.model.state_dict().items()}
        exact = _f64_logits(kind, sd, s.get_graph_positions(), s.get_depots()[:, 0], s.get_demands()[:, :, 0], ref_tape)
        _check_trained_logits(got, ref_logits, exact, 5.0, (kind, tag))
