"""CPU-only: the C MT19937 legacy-stream generator (csrc/mt19937_legacy.cu) is bit-exact with numpy's
global RandomState on the call sequence the reference makes (gym_vrp/graph/vrp_graph.py:29,34,42; gym_vrp/envs/tsp.py:48,55;
agents/random_agent.py:35), continues / hands back numpy's state, and feeds VRPNetwork / RandomAgent."""
import ctypes as C

import numpy as np
import pytest


def _numpy_instances(G, N, D):
    """The reference's per-graph numpy calls (vrp_graph.py:27-45)."""
    xy = np.empty((G, N, 2))
    dep = np.empty((G, D), np.int64)
    dem = np.empty((G, N))
    Cc = 0.2449 * N + 26.12
    for g in range(G):
        xy[g] = np.random.rand(N, 2)
        dep[g] = np.random.choice(N, size=D, replace=False)
        d = np.random.uniform(low=1, high=10, size=(N, 1)) / Cc
        d[dep[g]] = 0
        dem[g] = d[:, 0]
    return xy, dep, dem


@pytest.mark.parametrize("G,N,D,seed", [(100000, 20, 1, 1234), (20000, 50, 1, 69), (3000, 100, 1, 7), (2000, 10, 5, 3),
                                        (500, 128, 1, 0), (64, 2, 1, 4294967295)])
def test_instances_bit_exact_with_numpy(G, N, D, seed):
    from vrpx import legacy_stream

    np.random.seed(seed)
    draw_ref = np.random.choice(G, min(3, G), replace=False)
    ref = _numpy_instances(G, N, D)
    tail_ref = np.random.rand(5)
    np.random.seed(seed)
    draw = legacy_stream.permutation_head(G, min(3, G))
    got = legacy_stream.draw_instances(G, N, D)
    tail = np.random.rand(5)   # numpy continues where the C generator stopped
    assert np.array_equal(draw, draw_ref)
    for a, b in zip(got, ref):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    assert np.array_equal(tail, tail_ref)


def test_seed_matches_numpy_seeding():
    import vrpx

    L = vrpx.lib()
    for seed in (0, 1, 69, 1234, 2 ** 32 - 1):
        key = np.empty(624, np.uint32)
        pos = C.c_int32(0)
        assert L.vrpx_mt19937_seed(seed, key.ctypes.data_as(C.c_void_p), C.cast(C.byref(pos), C.c_void_p)) == 0
        np.random.seed(seed)
        _, k2, p2, _, _ = np.random.get_state()
        assert np.array_equal(key, k2) and pos.value == p2


def test_random_actions_bit_exact_with_numpy():
    from vrpx import legacy_stream

    rs = np.random.RandomState(1)
    for B, N in ((256, 20), (1000, 50), (17, 128)):
        mask = (rs.rand(B, N) < 0.6).astype(np.float64)
        mask[np.arange(B), rs.randint(0, N, B)] = 0      # at least one feasible node
        mask[0] = 1
        mask[0, N // 2] = 0                               # a single feasible node consumes no random word
        np.random.seed(5)
        ref = np.array([np.random.choice(np.flatnonzero(mask[i] == 0), 1)[0] for i in range(B)])
        t_ref = np.random.rand()
        np.random.seed(5)
        got = legacy_stream.random_actions(mask)
        assert np.array_equal(got, ref) and np.random.rand() == t_ref
    import vrpx

    with pytest.raises(vrpx.VrpxError):
        legacy_stream.random_actions(np.ones((2, 4)))


def test_network_and_oracle_streams_agree():
    """VRPNetwork (product host store, C generator) == oracle draw_instances (numpy calls) incl. a continued stream."""
    from gym_vrp.graph.vrp_network import VRPNetwork
    from oracle.env_oracle import draw_instances

    np.random.seed(2468)
    a1 = VRPNetwork(300, 30, 1)
    a2 = VRPNetwork(300, 30, 1)   # the reset() case: no reseed in between (tsp.py:150-160)
    np.random.seed(2468)
    for net in (a1, a2):
        xy, depot, demand = draw_instances(300, 30)
        assert np.array_equal(net.get_graph_positions(), xy)
        assert np.array_equal(net.get_depots()[:, 0], depot)
        assert np.array_equal(net.get_demands()[:, :, 0], demand)
