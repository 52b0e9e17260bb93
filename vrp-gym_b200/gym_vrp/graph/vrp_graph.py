"""VRPGraph — one routing instance as a *view* onto struct-of-arrays storage.

Reference: gym_vrp/graph/vrp_graph.py:5-146 builds one `networkx.complete_graph` per instance (372 KB at
N=50).  Here an instance is row `g` of the owning store's arrays (coordinates f64 (N,2), depot flags,
demand f64 (N,)); the attribute protocol the reference's callers use —

    graph.nodes[n]["coordinates"] = xy            (tests/test_env.py:31-36 via nx.set_node_attributes)
    nx.get_node_attributes(graph.graph, "depot")  (tests/test_graph.py:20)

— is kept through small mapping views that write through to the arrays (and bump the store's version so a
device-resident environment re-uploads).  A real networkx graph is only materialised on demand (`.graph`).
"""
from __future__ import annotations

import numpy as np


class _Store:
    """SoA storage shared by a VRPNetwork and its graph views."""

    def __init__(self, num_graphs: int, num_nodes: int, num_depots: int):
        self.xy = np.zeros((num_graphs, num_nodes, 2), dtype=np.float64)
        self.depots = np.zeros((num_graphs, num_depots), dtype=int)
        self.demand = np.zeros((num_graphs, num_nodes), dtype=np.float64)
        self.edges = [None] * num_graphs  # per graph: list of (source, target) marked as visited (drawing only)
        self.version = 0

    def draw_all(self):
        """Sample every instance from the legacy global numpy stream in the reference's order — per graph
        vrp_graph.py:29 rand(N,2) -> :34 choice(N, depots, replace=False) -> :42 uniform(1,10,(N,1))/C, one graph
        after the other (vrp_network.py:41-42) — in ONE call into the C generator (csrc/mt19937_legacy.cu) that
        continues numpy's global state, instead of three numpy calls per graph."""
        from vrpx import legacy_stream

        depots = np.empty(self.depots.shape, np.int64)
        legacy_stream.draw_instances(self.xy.shape[0], self.xy.shape[1], self.depots.shape[1], out=(self.xy, depots, self.demand))
        self.depots[:] = depots


class _NodeAttrs:
    """dict-like attribute view of one node: coordinates / depot / demand / node_color."""

    __slots__ = ("_s", "_g", "_n")
    _KEYS = ("coordinates", "depot", "demand", "node_color")

    def __init__(self, store, g, n):
        self._s, self._g, self._n = store, g, n

    def __getitem__(self, key):
        s, g, n = self._s, self._g, self._n
        if key == "coordinates":
            return s.xy[g, n].copy()
        if key == "depot":
            return 1.0 if n in s.depots[g] else 0.0
        if key == "demand":
            return np.array([s.demand[g, n]])
        if key == "node_color":
            return "red" if n in s.depots[g] else "black"
        raise KeyError(key)

    def __setitem__(self, key, value):
        s, g, n = self._s, self._g, self._n
        if key == "coordinates":
            s.xy[g, n] = np.asarray(value, dtype=np.float64)
        elif key == "demand":
            s.demand[g, n] = float(np.asarray(value).reshape(-1)[0])
        elif key == "depot":
            raise KeyError("depot flags are fixed at construction")
        elif key == "node_color":
            return
        else:
            raise KeyError(key)
        s.version += 1

    def keys(self):
        return self._KEYS

    def __iter__(self):
        return iter(self._KEYS)

    def __len__(self):
        return len(self._KEYS)

    def __contains__(self, key):
        return key in self._KEYS

    def items(self):
        return [(k, self[k]) for k in self._KEYS]

    def get(self, key, default=None):
        return self[key] if key in self._KEYS else default


class _NodeView:
    """`graph.nodes`: len / iteration over (n, attrs) / indexing by node id."""

    def __init__(self, store, g):
        self._s, self._g = store, g

    def __len__(self):
        return self._s.xy.shape[1]

    def __getitem__(self, n):
        n = int(n)
        if not 0 <= n < len(self):
            raise KeyError(n)
        return _NodeAttrs(self._s, self._g, n)

    def __iter__(self):
        return ((n, _NodeAttrs(self._s, self._g, n)) for n in range(len(self)))

    def __contains__(self, n):
        return 0 <= int(n) < len(self)


class VRPGraph:
    """One instance.  Constructed standalone it owns a one-graph store and samples itself from the global
    numpy stream exactly like the reference constructor (vrp_graph.py:9-47)."""

    def __init__(self, num_nodes: int, num_depots: int, plot_demand: bool = False, *, _store=None, _index=0):
        self.num_nodes = num_nodes
        self.num_depots = num_depots
        self.plot_demand = plot_demand
        self.offset = np.array([0, 0.065])
        if _store is None:
            _store = _Store(1, num_nodes, num_depots)
            _store.draw_all()
            _index = 0
        self._s, self._g = _store, _index

    # ---- reference attribute surface (vrp_graph.py:113-135)
    @property
    def depots(self) -> np.ndarray:
        return self._s.depots[self._g]

    @property
    def demand(self) -> np.ndarray:
        return self._s.demand[self._g][:, None].copy()

    @property
    def node_positions(self) -> np.ndarray:
        return self._s.xy[self._g].copy()

    @property
    def nodes(self):
        return _NodeView(self._s, self._g)

    @property
    def edges(self):
        vis = set(map(tuple, self._s.edges[self._g] or []))
        n = self.num_nodes
        return [(i, j, {"visited": (i, j) in vis or (j, i) in vis}) for i in range(n) for j in range(i + 1, n)]

    @property
    def graph(self):
        """A networkx materialisation (inspection / drawing only, built on demand)."""
        import networkx as nx

        G = nx.complete_graph(self.num_nodes)
        for n, attrs in self.nodes:
            for k in ("coordinates", "depot", "demand", "node_color"):
                G.nodes[n][k] = attrs[k]
        nx.set_edge_attributes(G, False, "visited")
        for (i, j) in self._s.edges[self._g] or []:
            G.edges[i, j]["visited"] = True
        return G

    def euclid_distance(self, node1_idx: int, node2_idx: int) -> float:
        """Host-side single-pair distance (vrp_graph.py:137-146); the batched path is the CUDA step kernel."""
        xy = self._s.xy[self._g]
        return np.linalg.norm(xy[node1_idx] - xy[node2_idx])

    def visit_edge(self, source_node: int, target_node: int) -> None:
        """Mark an edge for drawing (vrp_graph.py:98-111); self loops are ignored."""
        if source_node == target_node:
            return
        if self._s.edges[self._g] is None:
            self._s.edges[self._g] = []
        self._s.edges[self._g].append((int(source_node), int(target_node)))

    def set_default_node_attributes(self):
        self._s.edges[self._g] = None

    def draw(self, ax):
        """Rendering is outside the accelerated path; drawn from a networkx copy."""
        import networkx as nx

        G = self.graph
        pos = nx.get_node_attributes(G, "coordinates")
        colors = list(nx.get_node_attributes(G, "node_color").values())
        nx.draw_networkx_nodes(G, pos, node_color=colors, ax=ax, node_size=100)
        edges = [e for e in G.edges(data=True) if e[2]["visited"]]
        nx.draw_networkx_edges(G, pos, alpha=0.5, edgelist=edges, edge_color="red", ax=ax, width=1.5)
        if self.plot_demand:
            label_pos = {k: (v + self.offset) for k, v in pos.items()}
            labels = {k: np.round(v, 2)[0] for k, v in nx.get_node_attributes(G, "demand").items()}
            nx.draw_networkx_labels(G, label_pos, labels=labels, ax=ax)
