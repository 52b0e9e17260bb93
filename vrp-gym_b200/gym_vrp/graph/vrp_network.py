"""VRPNetwork — a batch of routing instances in struct-of-arrays form.

Reference: gym_vrp/graph/vrp_network.py:8-169 keeps a Python list of networkx graphs and loops over it in
every accessor.  Here the batch is three dense arrays (coordinates (B,N,2) f64, depots (B,D) int, demand
(B,N) f64) that upload to the GPU in one copy; `graphs[i]` are write-through views (vrp_graph.py).
The sampling order of the legacy global numpy stream is the reference's (per graph: coordinates, depot,
demand), so equal seeds give equal instances.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .vrp_graph import VRPGraph, _Store


class VRPNetwork:
    def __init__(self, num_graphs: int, num_nodes: int, num_depots: int, plot_demand: bool = False,
                 *, _sample: bool = True):
        assert num_nodes >= num_depots, "Number of depots should be lower than number of depots"
        self.num_nodes = num_nodes
        self.num_depots = num_depots
        self.num_graphs = num_graphs
        self.plot_demand = plot_demand
        self._store = _Store(num_graphs, num_nodes, num_depots)
        if _sample:
            self._store.draw_all()  # vrp_network.py:41-42 — one instance after the other, same stream order
        self.graphs: List[VRPGraph] = _GraphList(self)

    # ---- bulk accessors (vrp_network.py:80-108, :154-169): array views instead of per-graph loops
    def get_graph_positions(self) -> np.ndarray:
        return self._store.xy.copy()

    def get_depots(self) -> np.ndarray:
        return self._store.depots.copy()

    def get_demands(self) -> np.ndarray:
        return self._store.demand[:, :, None].copy()

    @property
    def version(self) -> int:
        return self._store.version

    # ---- host-side distance helpers (vrp_network.py:44-78); the batched hot path is vrpx_env_step
    def get_distance(self, graph_idx: int, node_idx_1: int, node_idx_2: int) -> float:
        return self.graphs[graph_idx].euclid_distance(node_idx_1, node_idx_2)

    def get_distances(self, paths) -> np.ndarray:
        paths = np.asarray(paths).astype(int)
        ar = np.arange(self.num_graphs)
        d = self._store.xy[ar, paths[:, 0]] - self._store.xy[ar, paths[:, 1]]
        return np.array([np.linalg.norm(row) for row in d])

    def visit_edges(self, transition_matrix: np.ndarray, only=None) -> None:
        """Record traversed edges for drawing (vrp_network.py:143-152).  `only` restricts the bookkeeping to
        the graphs that can ever be rendered (the env's draw_idxs)."""
        idxs = range(len(transition_matrix)) if only is None else only
        for i in idxs:
            row = transition_matrix[i]
            self.graphs[i].visit_edge(int(row[0]), int(row[1]))

    def draw(self, graph_idxs: np.ndarray):
        """Matplotlib grid of selected graphs (vrp_network.py:110-141) — outside the accelerated path."""
        import matplotlib.pyplot as plt

        num_columns = min(len(graph_idxs), 3)
        num_rows = np.ceil(len(graph_idxs) / num_columns).astype(int)
        plt.clf()
        fig = plt.figure(figsize=(5 * num_columns, 5 * num_rows))
        for n, graph_idx in enumerate(graph_idxs):
            ax = plt.subplot(num_rows, num_columns, n + 1)
            self.graphs[graph_idx].draw(ax=ax)
        plt.show()
        fig.canvas.draw()
        data = np.frombuffer(fig.canvas.tostring_rgb(), dtype=np.uint8)
        return data.reshape(fig.canvas.get_width_height()[::-1] + (3,))


class _GraphList:
    """Indexable, sized, iterable collection of VRPGraph views (created lazily)."""

    def __init__(self, net: VRPNetwork):
        self._net = net

    def __len__(self):
        return self._net.num_graphs

    def __getitem__(self, i):
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        n = self._net
        return VRPGraph(n.num_nodes, n.num_depots, n.plot_demand, _store=n._store, _index=i)

    def __iter__(self):
        return (self[i] for i in range(len(self)))
