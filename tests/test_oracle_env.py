"""Pin the env oracle (oracle/env_oracle.py) against the reference's own artefacts:
  * the 6,912 Random-Agent costs of reference reproduction_log/*.csv (tests/golden/random_agent_costs.npz);
  * transition tapes recorded from the unmodified reference (tests/golden/env_tapes.npz);
  * known answers of reference tests/test_env.py, tests/test_agent.py:69 and SURVEY App. C."""
import json
import os

import numpy as np
import pytest

from oracle.env_oracle import EnvOracle, random_agent_rollout, seeded_env_instances


def _keys(z):
    return sorted({k.split("/")[0] for k in z.files})


def test_env_tapes_bit_exact(golden_dir):
    z = np.load(os.path.join(golden_dir, "env_tapes.npz"))
    for key in _keys(z):
        kind, N, B, seed = key.split("_")
        N, B, seed = int(N), int(B), int(seed)
        draw_idxs, xy, depot, demand = seeded_env_instances(N, B, min(3, B), seed)
        # instance stream: same arrays as the reference constructor
        assert np.array_equal(draw_idxs, z[key + "/draw_idxs"]), key
        assert np.array_equal(xy, z[key + "/xy"]) and np.array_equal(depot, z[key + "/depot"]), key
        assert np.array_equal(demand, z[key + "/demand"]), key
        env = EnvOracle(kind, xy, depot, demand)
        st = env.get_state()
        st = st[0] if kind == "irp" else st
        assert np.array_equal(st, z[key + "/state0"]), key
        for t, a in enumerate(z[key + "/actions"]):
            st, r, done, _ = env.step(a[:, None])
            if kind == "irp":
                st, load = st
                assert np.array_equal(load, z[key + "/load"][t]), (key, t)
            assert np.array_equal(env.visited.astype(np.uint8), z[key + "/visited"][t]), (key, t)
            assert np.array_equal(st[:, :, -1].astype(np.uint8), z[key + "/mask"][t]), (key, t)
            assert done == z[key + "/done"][t], (key, t)
            assert np.array_equal(r, z[key + "/reward"][t]), (key, t)  # f64 bit-exact (fma emulation)
        # reset(): the stream continues without reseeding (tsp.py:150-160)
        from oracle.env_oracle import draw_instances
        xy2, depot2, _ = draw_instances(B, N)
        assert np.array_equal(xy2, z[key + "/reset_xy"]) and np.array_equal(depot2, z[key + "/reset_depot"]), key


@pytest.mark.parametrize("kind", ["tsp", "vrp", "irp"])
@pytest.mark.parametrize("N", [20, 30, 40])
def test_random_agent_golden_csv(golden_dir, kind, N):
    """reproduction.py:32-48 path: Env(N,256,3,seed) -> RandomAgent(seed)(env); costs equal the published CSV rows."""
    z = np.load(os.path.join(golden_dir, "random_agent_costs.npz"))
    for seed in (1234, 2468, 2048):
        _, xy, depot, demand = seeded_env_instances(N, 256, 3, seed)
        env = EnvOracle(kind, xy, depot, demand)
        cost = -random_agent_rollout(env, seed).numpy()
        gold = z[f"{kind}_{N}_{seed}"]
        assert np.array_equal(cost, np.float32(gold)), (kind, N, seed)


def test_reference_known_answers(golden_dir):
    ka = json.load(open(os.path.join(golden_dir, "known_answers.json")))
    # reference tests/test_env.py: seed 69, VRPEnv(3, 2, 2) -> depots [1, 2]
    np.random.seed(69)
    draw = np.random.choice(2, 2, replace=False)
    from oracle.env_oracle import draw_instances
    xy, depot, demand = draw_instances(2, 3)
    assert depot.tolist() == ka["seed69_vrp_3_2_depots"] and draw.tolist() == ka["seed69_vrp_3_2_draw_idxs"]
    # test_step (tests/test_env.py:44-48): equilateral triangles of side 1 and 4, action 2 -> rewards [-1, 0]
    y = np.sqrt(3) / 2
    xy = np.array([[[0, 0], [1, 0], [0.5, y]], [[0, 0], [4, 0], [2, 4 * y]]], dtype=np.float64)
    env = EnvOracle("vrp", xy, depot, demand)
    st = env.get_state()
    assert st.shape == (2, 3, 4) and st[:, :, 2].sum() == 2
    st, r, _, _ = env.step(np.array([2, 2])[:, None])
    assert np.allclose(r, [-1, 0]) and st[0, 2, 3] == 1 and st[1, 2, 3] == 1
    # reference tests/test_agent.py:57-69 (session seed 69; env seed 69; RandomAgent() reseeds 69)
    _, xy, depot, demand = seeded_env_instances(8, 2, 1, 69)
    loss = random_agent_rollout(EnvOracle("vrp", xy, depot, demand), 69)
    assert np.isclose(loss.mean().item(), ka["test_random_agent_mean"])
    # SURVEY App. C
    draw, xy, depot, _ = seeded_env_instances(20, 256, 3, 1234)
    assert draw.tolist() == ka["tsp_20_256_1234_draw_idxs"]
    assert depot[:8].tolist() == ka["tsp_20_256_1234_depots8"]
    assert xy[0, 0].tolist() == ka["tsp_20_256_1234_xy00"]
