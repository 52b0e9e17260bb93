"""GPU parity of the environment kernels (csrc/env.cu, env_rules.cuh) through the reference-shaped Python
surface (gym_vrp.envs.*), against the numpy oracle and the golden fixtures.  Integer state (visited, mask, done,
actions) and the f64 load/reward are required to be BIT-EXACT."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ENVS = {}


def _envs():
    from gym_vrp.envs import IRPEnv, TSPEnv, VRPEnv

    return {"tsp": TSPEnv, "vrp": VRPEnv, "irp": IRPEnv}


def _keys(z):
    return sorted({k.split("/")[0] for k in z.files})


def test_transition_tapes_bit_exact(golden_dir):
    """Same seeds -> same instances (numpy legacy stream on the host), same transitions as the reference."""
    z = np.load(os.path.join(golden_dir, "env_tapes.npz"))
    for key in _keys(z):
        kind, N, B, seed = key.split("_")
        N, B, seed = int(N), int(B), int(seed)
        env = _envs()[kind](N, B, min(3, B), seed)
        assert np.array_equal(env.draw_idxs, z[key + "/draw_idxs"])
        assert np.array_equal(env.sampler.get_graph_positions(), z[key + "/xy"])
        assert np.array_equal(env.depots[:, 0], z[key + "/depot"])
        st = env.get_state()
        st = st[0] if kind == "irp" else st
        assert st.dtype == np.float64 and np.array_equal(st, z[key + "/state0"]), key
        for t, a in enumerate(z[key + "/actions"]):
            st, r, done, info = env.step(a[:, None])
            if kind == "irp":
                st, load = st
                assert np.array_equal(load, z[key + "/load"][t]), (key, t)
            assert info is None and isinstance(done, (bool, np.bool_))
            assert np.array_equal(env.visited.astype(np.uint8), z[key + "/visited"][t]), (key, t)
            assert np.array_equal(st[:, :, -1].astype(np.uint8), z[key + "/mask"][t]), (key, t)
            assert np.array_equal(env.generate_mask().astype(np.uint8), z[key + "/mask"][t]), (key, t)
            assert bool(done) == bool(z[key + "/done"][t]), (key, t)
            assert r.dtype == np.float64 and np.array_equal(r, z[key + "/reward"][t]), (key, t)
            assert np.array_equal(env.current_location[:, 0], a)
        assert env.step_count == len(z[key + "/actions"])
        env.reset()  # continues the stream, no reseed (tsp.py:150-160)
        assert env.step_count == 0
        assert np.array_equal(env.sampler.get_graph_positions(), z[key + "/reset_xy"]), key
        assert np.array_equal(env.depots[:, 0], z[key + "/reset_depot"]), key


@pytest.mark.parametrize("kind,N,seed", [("tsp", 20, 1234), ("vrp", 20, 2468), ("irp", 20, 2048), ("tsp", 40, 2048),
                                         ("vrp", 30, 1234), ("irp", 40, 1234)])
def test_random_agent_reproduces_published_csv(golden_dir, kind, N, seed):
    """reproduction.py:32-48 on the CUDA env: the per-instance Random-Agent costs of the reference's
    reproduction_log/*.csv (float32-exact)."""
    from copy import deepcopy

    from agents import RandomAgent

    z = np.load(os.path.join(golden_dir, "random_agent_costs.npz"))
    env = _envs()[kind](num_nodes=N, batch_size=256, num_draw=3, seed=seed)
    env_r = deepcopy(env)
    loss = RandomAgent(seed=seed)(env_r)
    assert np.array_equal(-loss.numpy(), np.float32(z[f"{kind}_{N}_{seed}"]))
    assert env.step_count == 0 and env_r.step_count > 0  # deepcopy gave an independent device snapshot


def test_reference_env_tests():
    """reference tests/test_env.py: coordinates written through nx.set_node_attributes are seen by step()."""
    import math

    import networkx as nx

    from gym_vrp.envs import VRPEnv

    np.random.seed(69)
    env = VRPEnv(3, 2, 2)
    y = math.sqrt(3) / 2
    nx.set_node_attributes(env.sampler.graphs[0], {0: np.array([0, 0]), 1: np.array([1, 0]), 2: np.array([0.5, y])}, "coordinates")
    nx.set_node_attributes(env.sampler.graphs[1], {0: np.array([0, 0]), 1: np.array([4, 0]), 2: np.array([2, 4 * y])}, "coordinates")
    assert len(env.sampler.graphs) == 2 and len(env.sampler.graphs[0].nodes) == 3
    state = env.get_state()
    assert state.shape == (2, 3, 4) and np.sum(state[:, :, 2]) == 2
    state, reward, _, _ = env.step(np.array([2, 2])[:, None])
    assert np.allclose(reward, np.array([-1, 0]))
    assert state[0, 2, 3] == 1 and state[1, 2, 3] == 1


def test_assertions_match_reference():
    from gym_vrp.envs import TSPEnv

    with pytest.raises(AssertionError):
        TSPEnv(num_nodes=5, batch_size=2, num_draw=3)
    env = TSPEnv(num_nodes=5, batch_size=4, num_draw=1)
    with pytest.raises(AssertionError):
        env.step(np.zeros((3, 1), dtype=int))


@pytest.mark.parametrize("kind", ["tsp", "vrp", "irp"])
def test_large_random_walk_vs_oracle(kind):
    """Full-size style check at B=4096, N=100: random feasible walks, every transition compared with the oracle."""
    from oracle.env_oracle import EnvOracle

    B, N = 4096, 100
    env = _envs()[kind](N, B, 0, seed=5, instance_rng="philox")
    s = env.sampler
    xy, depot, demand = s.get_graph_positions(), s.get_depots()[:, 0], s.get_demands()[:, :, 0]
    assert xy.min() >= 0 and xy.max() < 1 and depot.min() >= 0 and depot.max() < N
    if kind == "irp":
        C = 0.2449 * N + 26.12
        assert demand[np.arange(B), depot].max() == 0 and demand.max() < 10 / C
    orc = EnvOracle(kind, xy, depot, demand)
    st_o = orc.get_state()
    st = env.get_state()
    if kind == "irp":
        st_o, st = st_o[0], st[0]
    assert np.array_equal(st, st_o)
    rs = np.random.RandomState(0)
    done, t = False, 0
    while not done:
        mask = st[:, :, -1]
        # vectorised random feasible choice
        score = rs.rand(B, N) - mask * 2
        a = score.argmax(1)
        st, r, done, _ = env.step(a[:, None])
        st_o, r_o, done_o, _ = orc.step(a[:, None])
        if kind == "irp":
            assert np.array_equal(st[1], st_o[1])
            st, st_o = st[0], st_o[0]
        assert np.array_equal(st, st_o) and np.array_equal(r, r_o) and bool(done) == bool(done_o), t
        t += 1
    assert t >= N - 1


def test_philox_generator_is_deterministic_and_shardable():
    from gym_vrp.envs import VRPEnv

    a = VRPEnv(30, 512, 0, seed=11, instance_rng="philox")
    b = VRPEnv(30, 256, 0, seed=11, instance_rng="philox", instance_offset=256)
    xa, xb = a.sampler.get_graph_positions(), b.sampler.get_graph_positions()
    assert np.array_equal(xa[256:], xb)  # instance id keyed counters: a shard equals the slice of the full batch
    assert np.array_equal(a.depots[256:], b.depots)
    hist = np.bincount(a.depots[:, 0], minlength=30)
    assert hist.min() > 0
