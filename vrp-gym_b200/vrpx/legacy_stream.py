"""The reference's seed-compatible instance stream on the C side of the ABI (include/vrpx.h, csrc/mt19937_legacy.cu).

The reference consumes numpy's legacy GLOBAL RandomState per graph (gym_vrp/graph/vrp_graph.py:29,34,42) and per
RandomAgent action (agents/random_agent.py:35).  Here the same words are drawn by C code working on a copy of numpy's
generator state, which is handed back afterwards — so `np.random.seed(s)`, these calls and any later `np.random.*` call
interleave exactly as in the reference (equal seeds -> equal instances, also after `reset()`), without three numpy
calls per graph or one per instance-step.  Host only: no GPU involved.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import vrpx


class _GlobalState:
    """np.random.get_state() on entry, np.random.set_state() with the advanced key/pos on exit."""

    def __enter__(self):
        name, key, pos, has_gauss, cached = np.random.get_state()
        assert name == "MT19937"
        self.key = np.ascontiguousarray(key, dtype=np.uint32)
        self.pos = C.c_int32(int(pos))
        self._rest = (int(has_gauss), float(cached))
        return self

    def __exit__(self, *exc):
        np.random.set_state(("MT19937", self.key, int(self.pos.value)) + self._rest)
        return False

    @property
    def args(self):
        return self.key.ctypes.data_as(C.c_void_p), C.cast(C.byref(self.pos), C.c_void_p)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def draw_instances(num_graphs: int, num_nodes: int, num_depots: int = 1, out=None):
    """`num_graphs` instances from the current global numpy stream, in the reference's order.  Returns
    (xy (G,N,2) f64, depots (G,D) int64, demand (G,N) f64); `out` = the same three arrays to fill in place."""
    if out is None:
        out = (np.empty((num_graphs, num_nodes, 2), np.float64), np.empty((num_graphs, num_depots), np.int64),
               np.empty((num_graphs, num_nodes), np.float64))
    xy, depots, demand = out
    assert xy.flags.c_contiguous and depots.flags.c_contiguous and demand.flags.c_contiguous
    assert xy.dtype == np.float64 and depots.dtype == np.int64 and demand.dtype == np.float64
    with _GlobalState() as st:
        vrpx.check(vrpx.lib().vrpx_mt19937_instances(*st.args, num_graphs, num_nodes, num_depots, _p(xy), _p(depots), _p(demand)))
    return xy, depots, demand


def permutation_head(n: int, k: int) -> np.ndarray:
    """np.random.choice(n, k, replace=False) on the global stream."""
    out = np.empty((k,), np.int64)
    with _GlobalState() as st:
        vrpx.check(vrpx.lib().vrpx_mt19937_permutation_head(*st.args, n, k, _p(out)))
    return out


def random_actions(mask: np.ndarray) -> np.ndarray:
    """Per instance np.random.choice(np.flatnonzero(mask[i] == 0), 1)[0] on the global stream; mask (B,N) of 0/1."""
    m = np.ascontiguousarray(mask, dtype=np.float64)
    out = np.empty((m.shape[0],), np.int64)
    with _GlobalState() as st:
        vrpx.check(vrpx.lib().vrpx_mt19937_random_actions(*st.args, _p(m), m.shape[0], m.shape[1], _p(out)))
    return out
