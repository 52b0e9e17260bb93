"""IRPModel / IRPAgent — inventory routing: three node features (x, y, demand) and the vehicle load in the
decoder context (reference agents/graph_irp_agent.py:12-170; context = _context_proj([graph, last, load]))."""
from __future__ import annotations

from .graph_vrp_agent import VRPAgent, VRPModel


class IRPModel(VRPModel):
    pass


class IRPAgent(VRPAgent):
    _NODE_DIM = 3
    _MODEL = IRPModel
