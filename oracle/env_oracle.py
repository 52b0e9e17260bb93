"""Vectorised numpy restatement of the reference environments and instance stream.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the
reference lines it follows (paths relative to the reference root).

Parity: PINNED against reference `reproduction_log/*.csv` Random-Agent rows and
transition tapes recorded from the unmodified reference (tests/golden/).
"""
from __future__ import annotations

import numpy as np

TSP, VRP, IRP = 0, 1, 2
KIND_BY_NAME = {"tsp": TSP, "vrp": VRP, "irp": IRP}


# --------------------------------------------------------------------------
# instance stream
# --------------------------------------------------------------------------
def draw_instances(num_graphs: int, num_nodes: int):
    """Draw `num_graphs` instances from the *current* legacy global numpy stream.

    Follows gym_vrp/graph/vrp_graph.py:27-45 (one graph) called num_graphs times
    in sequence by gym_vrp/graph/vrp_network.py:41-42: per graph
    rand(N,2) -> choice(N,1,replace=False) -> uniform(1,10,(N,1))/C, depot demand 0.
    Demand is drawn for every env kind (vrp_graph.py:41-45).
    """
    xy = np.empty((num_graphs, num_nodes, 2), dtype=np.float64)
    depot = np.empty((num_graphs,), dtype=np.int64)
    demand = np.empty((num_graphs, num_nodes), dtype=np.float64)
    C = 0.2449 * num_nodes + 26.12  # vrp_graph.py:41
    for g in range(num_graphs):
        xy[g] = np.random.rand(num_nodes, 2)  # vrp_graph.py:29
        depot[g] = np.random.choice(num_nodes, size=1, replace=False)[0]  # :34
        d = np.random.uniform(low=1, high=10, size=(num_nodes, 1)) / C  # :42
        d[depot[g]] = 0  # :43
        demand[g] = d[:, 0]
    return xy, depot, demand


def seeded_env_instances(num_nodes: int, batch_size: int, num_draw: int, seed: int):
    """Env constructor stream: seed, draw_idxs, instances (tsp.py:48,55,58)."""
    np.random.seed(seed)
    draw_idxs = np.random.choice(batch_size, num_draw, replace=False)
    xy, depot, demand = draw_instances(batch_size, num_nodes)
    return draw_idxs, xy, depot, demand


def _fma(a, b, c):
    """Correctly rounded a*b + c for float64 arrays (Dekker two-product + two-sum; no hardware FMA in numpy).
    Exact up to the final rounding except in astronomically rare double-rounding ties."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    split = 134217729.0  # 2^27 + 1
    p = a * b
    ta = split * a
    ah = ta - (ta - a)
    al = a - ah
    tb = split * b
    bh = tb - (tb - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl  # a*b == p + e exactly
    s = p + c
    bb = s - p
    t = (p - (s - bb)) + (c - bb)  # p + c == s + t exactly
    return s + (t + e)


# --------------------------------------------------------------------------
# environment
# --------------------------------------------------------------------------
class EnvOracle:
    """Batched TSP / VRP / IRP environment on plain arrays.

    State per instance: xy[N,2] f64, depot, demand[N] f64,
    visited[N] in {0,1}, cur, load f64 (IRP).
    """

    def __init__(self, kind, xy, depot, demand=None):
        self.kind = KIND_BY_NAME[kind] if isinstance(kind, str) else int(kind)
        self.xy = np.asarray(xy, dtype=np.float64)
        self.B, self.N, _ = self.xy.shape
        self.depot = np.asarray(depot, dtype=np.int64).reshape(self.B)
        if demand is None:
            demand = np.zeros((self.B, self.N))
        self.demand = np.asarray(demand, dtype=np.float64).reshape(self.B, self.N)
        self._ar = np.arange(self.B)
        self.reset_episode()

    # tsp.py:158-160,167-174 ; irp.py:47,183-185
    def reset_episode(self):
        self.visited = np.zeros((self.B, self.N), dtype=np.float64)
        self.cur = self.depot.copy()
        self.load = np.ones((self.B,), dtype=np.float64)
        self.step_count = 0

    # tsp.py:131-148 | vrp.py:13-37 | irp.py:126-155
    def generate_mask(self):
        at_depot = self.cur == self.depot
        ar = self._ar
        # R1: disallow staying on the depot (tsp.py:141-142)
        self.visited[ar[at_depot], self.depot[at_depot]] = 1
        if self.kind != TSP:
            # R2: depot re-visitable when away from it (vrp.py:28-31, irp.py:141-144)
            self.visited[ar[~at_depot], self.depot[~at_depot]] = 0
        # R3: solved graphs may idle on the depot (tsp.py:145-146)
        done = np.all(self.visited == 1, axis=1)
        self.visited[ar[done], self.depot[done]] = 0
        if self.kind != IRP:
            return self.visited  # same storage, no copy (tsp.py:148)
        # R4 (irp.py:151-155): float64 compare demand - load > 0
        mask = self.visited.copy()
        mask[(self.demand - self.load[:, None]) > 0] = 1
        return mask

    # tsp.py:103-104
    def is_done(self):
        return bool(np.all(self.visited == 1))

    # tsp.py:106-129 ; irp.py:101-124
    def get_state(self):
        mask = self.generate_mask()
        is_depot = np.zeros((self.B, self.N))
        is_depot[self._ar, self.depot] = 1
        if self.kind == IRP:
            st = np.dstack([self.xy, self.demand[:, :, None], is_depot, mask])
            return st, self.load
        return np.dstack([self.xy, is_depot, mask])

    # tsp.py:60-101 ; irp.py:49-99
    def step(self, actions, observe=True):
        """observe=False skips building the (B,N,4|5) observation (the mask rules still run, as the reference's
        get_state call inside step would run them): for replaying long tapes at large batches in tests."""
        a = np.asarray(actions).reshape(self.B).astype(np.int64)
        ar = self._ar
        self.step_count += 1
        self.visited[ar, a] = 1  # tsp.py:86
        # vrp_graph.py:137-146: np.linalg.norm(p - q) = sqrt(ddot(d, d)); the BLAS ddot evaluates
        # fma(dy, dy, dx*dx) (verified bit-for-bit on the recorded reference rewards).
        d = self.xy[ar, self.cur] - self.xy[ar, a]
        dist = np.sqrt(_fma(d[:, 1], d[:, 1], d[:, 0] * d[:, 0]))
        if self.kind == IRP:
            self.load = self.load - self.demand[ar, a]  # irp.py:85
            self.load[a == self.depot] = 1  # irp.py:86
        self.cur = a  # tsp.py:90
        done = self.is_done()  # BEFORE the mask rules (tsp.py:95)
        if not observe:
            self.generate_mask()
            return None, -dist, done, None
        state = self.get_state()  # applies MASK() (tsp.py:97)
        return state, -dist, done, None


def random_agent_rollout(env: EnvOracle, seed: int):
    """agents/random_agent.py:11-41: reseeds the global stream, then per step and per
    instance `np.random.choice(feasible, 1)`; accumulates f32(reward) in step order."""
    import torch

    np.random.seed(seed)  # random_agent.py:13
    state = env.get_state()
    if isinstance(state, tuple):
        state = state[0]
    acc = torch.zeros(size=(state.shape[0],))
    done = False
    while not done:
        if isinstance(state, tuple):
            state = state[0]
        actions = []
        for i in range(state.shape[0]):
            pos = np.argwhere(state[i, :, -1] == 0).flatten()
            actions.append(np.random.choice(pos, 1)[0])
        state, loss, done, _ = env.step(np.array(actions)[:, None])
        acc += torch.tensor(loss, dtype=torch.float)
    return acc
