"""RandomAgent — uniformly random feasible actions (reference agents/random_agent.py:6-41).

The action stream is the legacy global numpy generator consumed instance by instance
(`np.random.choice(feasible, 1)`), which is inherently sequential; it therefore stays on the host (in C, csrc/mt19937_legacy.cu, continuing numpy's
own generator state) so that equal seeds give the reference's exact tours.  Transitions, distances and masks come from the CUDA env."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from vrpx import legacy_stream


class RandomAgent(nn.Module):
    def __init__(self, seed: int = 69):
        super().__init__()
        np.random.seed(seed)

    def forward(self, env) -> torch.Tensor:
        state = env.get_state()
        if isinstance(state, tuple):  # IRPEnv
            state = state[0]
        done = False
        acc_loss = torch.zeros(size=(state.shape[0],))
        while not done:
            if isinstance(state, tuple):
                state = state[0]
            # per instance np.random.choice(feasible, 1) on the global stream (reference :33-35), drawn by the C
            # generator that continues numpy's state (vrpx/legacy_stream.py) instead of a Python loop over the batch
            actions = legacy_stream.random_actions(state[:, :, -1])
            state, loss, done, _ = env.step(actions[:, None])
            acc_loss += torch.tensor(loss, dtype=torch.float)  # f32 accumulation, as the reference
        return acc_loss
