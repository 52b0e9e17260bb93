"""CPU-only checks: the C-ABI library loads and exports every symbol include/vrpx.h declares (no compute calls),
the product refuses to run without a GPU (no CPU fallback), and the host-side instance store mirrors the
reference's VRPGraph / VRPNetwork behaviour (reference tests/test_graph.py)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vrpx.h")).read()
    return sorted(set(re.findall(r"VRPX_API\s+[\w\s\*]+?\b(vrpx_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import vrpx

    L = vrpx.lib()
    names = _declared_symbols()
    assert len(names) >= 14, names
    for n in names:
        assert hasattr(L, n), f"libvrpx.so does not export {n}"
    assert L.vrpx_abi_version() == vrpx.ABI_VERSION
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'vrpx.h')).read()
    assert f'#define VRPX_ABI_VERSION {vrpx.ABI_VERSION}' in header
    assert L.vrpx_launch_count() == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import vrpx
    from gym_vrp.envs import TSPEnv

    assert vrpx.lib().vrpx_device_check(0) < 0
    assert b"no CUDA device" in vrpx.lib().vrpx_last_error() or b"device" in vrpx.lib().vrpx_last_error()
    with pytest.raises(vrpx.VrpxError):
        vrpx.require_device()
    with pytest.raises(RuntimeError):
        TSPEnv(num_nodes=5, batch_size=2, num_draw=1)
    from agents import GraphEncoder

    with pytest.raises(RuntimeError):
        GraphEncoder(node_input_dim=2)(torch.zeros(2, 5, 2))


# ---- reference tests/test_graph.py on the SoA-backed host classes
def test_vrp_graph_init():
    from gym_vrp.graph.vrp_graph import VRPGraph
    import networkx as nx

    np.random.seed(69)
    graph = VRPGraph(10, 5)
    depots = list(nx.get_node_attributes(graph.graph, "depot").values())
    assert len(graph.nodes) == 10
    assert depots.count(1) == 5


def test_vrp_euclid_dist_write_through():
    from gym_vrp.graph.vrp_graph import VRPGraph
    import networkx as nx

    np.random.seed(69)
    graph = VRPGraph(2, 1)
    nx.set_node_attributes(graph, {0: np.array([2, -1]), 1: np.array([-2, 2])}, "coordinates")
    assert graph.euclid_distance(0, 1) == 5


def test_vrp_network_init_and_stream_order():
    from gym_vrp.graph.vrp_network import VRPNetwork
    from oracle.env_oracle import draw_instances

    np.random.seed(69)
    net = VRPNetwork(num_graphs=10, num_nodes=10, num_depots=2)
    assert len(net.graphs) == 10
    depots = net.get_depots()
    assert depots.shape == (10, 2) and np.unique(depots, axis=0).shape[0] > 1
    # single-depot stream equals the oracle's restatement of the reference order
    np.random.seed(5)
    net = VRPNetwork(7, 13, 1)
    np.random.seed(5)
    xy, depot, demand = draw_instances(7, 13)
    assert np.array_equal(net.get_graph_positions(), xy)
    assert np.array_equal(net.get_depots()[:, 0], depot)
    assert np.array_equal(net.get_demands()[:, :, 0], demand)
    v0 = net.version
    net.graphs[3].nodes[2]["coordinates"] = np.array([0.25, 0.75])
    assert net.version == v0 + 1 and np.array_equal(net.get_graph_positions()[3, 2], [0.25, 0.75])
    assert net.get_demands().shape == (7, 13, 1)
    with pytest.raises(AssertionError):
        VRPNetwork(2, 3, 4)


def test_state_dict_layout_matches_reference_keys():
    """67 tensors (TSP) / 69 (VRP, IRP) with the reference's names and shapes."""
    from agents import IRPAgent, TSPAgent, VRPAgent

    sd = TSPAgent(seed=1).model.state_dict()
    assert len(sd) == 67
    assert sd["encoder.attention_layers.2.attention_layer.in_proj_weight"].shape == (384, 128)
    assert sd["encoder.attention_layers.0.ff.0.weight"].shape == (512, 128)
    assert sd["encoder.attention_layers.1.bn2.norm.running_var"].shape == (128,)
    assert sd["decoder._first_node"].shape == (1, 1, 128)
    assert sd["decoder.attention.q_proj_weight"].shape == (384, 384)
    assert sd["decoder.attention.k_proj_weight"].shape == (384, 128)
    assert sd["decoder.attention.in_proj_bias"].shape == (1152,)
    assert sd["decoder._context_proj.weight"].shape == (384, 257)
    assert sd["decoder._att_output.weight"].shape == (128, 384)
    for A, f in ((VRPAgent, 2), (IRPAgent, 3)):
        sd = A(seed=1).model.state_dict()
        assert len(sd) == 69
        assert sd["encoder.node_embed.weight"].shape == (128, f)
        assert sd["encoder.depot_embed.weight"].shape == (128, 2)
