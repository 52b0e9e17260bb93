"""IRPEnv — inventory routing: a vehicle of capacity 1 serves node demands and refills at the depot
(reference gym_vrp/envs/irp.py:10-185).

State is `(graph_state (B,N,5) = [x, y, demand, is_depot, mask], load (B,))`.  The f64 load update
(`load -= demand[a]`, refill on the depot, irp.py:80-86) and the f64 mask rule `demand - load > 0`
(irp.py:151-155) run inside the CUDA transition (csrc/env_rules.cuh), bit-exact with numpy.
"""
from __future__ import annotations

import numpy as np

import vrpx

from .tsp import TSPEnv


class IRPEnv(TSPEnv):
    metadata = {"render.modes": ["human", "rgb_array"]}

    _KIND = vrpx.IRP
    _PLOT_DEMAND = True
    _STATE_COLS = 5

    def __init__(self, num_nodes: int = 32, batch_size: int = 128, num_draw: int = 6, seed: int = 69, **kw):
        super().__init__(num_nodes=num_nodes, batch_size=batch_size, num_draw=num_draw, seed=seed, **kw)

    @property
    def load(self) -> np.ndarray:
        """Vehicle load (B,) f64 — a host copy of the device array."""
        return self._load.cpu().numpy()

    @property
    def demands(self) -> np.ndarray:
        return self._demand.cpu().numpy()[:, :, None]

    def _wrap_state(self, state_np):
        return state_np, self.load
