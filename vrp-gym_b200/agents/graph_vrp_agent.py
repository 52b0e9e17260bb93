"""VRPModel / VRPAgent — as the TSP pair, with the depot-aware encoder (reference
agents/graph_vrp_agent.py:10-148).  The depot row is embedded by `encoder.depot_embed`; the reference reads
the depot one-hot from state column 3, which equals the depot flag at reset (:67)."""
from __future__ import annotations

from .graph_encoder import GraphDemandEncoder
from .graph_tsp_agent import TSPAgent, TSPModel


class VRPModel(TSPModel):
    _USES_DEPOT_EMBED = True

    def __init__(self, depot_dim: int, node_dim: int, emb_dim: int, hidden_dim: int, num_attention_layers: int,
                 num_heads: int):
        super().__init__(node_dim=node_dim, emb_dim=emb_dim, hidden_dim=hidden_dim,
                         num_attention_layers=num_attention_layers, num_heads=num_heads)
        self.encoder = GraphDemandEncoder(depot_input_dim=depot_dim, node_input_dim=node_dim, embedding_dim=emb_dim,
                                          hidden_dim=hidden_dim, num_attention_layers=num_attention_layers,
                                          num_heads=num_heads)


class VRPAgent(TSPAgent):
    _NODE_DIM = 2

    def __init__(self, depot_dim: int = 2, node_dim: int = None, emb_dim: int = 128, hidden_dim: int = 512,
                 num_attention_layers: int = 3, num_heads: int = 8, lr: float = 1e-4, csv_path: str = "loss_log.csv",
                 seed=69):
        node_dim = self._NODE_DIM if node_dim is None else node_dim
        # The reference first builds the TSP pair in the parent constructor, then the demand-aware pair
        # (graph_vrp_agent.py:108-143); the order is kept because it fixes the seeded initial weights.
        super().__init__(node_dim=node_dim, emb_dim=emb_dim, hidden_dim=hidden_dim,
                         num_attention_layers=num_attention_layers, num_heads=num_heads, lr=lr, csv_path=csv_path,
                         seed=seed)
        kw = dict(depot_dim=depot_dim, node_dim=node_dim, emb_dim=emb_dim, hidden_dim=hidden_dim,
                  num_attention_layers=num_attention_layers, num_heads=num_heads)
        self.model = self._MODEL(**kw).to(self.device)
        self.target_model = self._MODEL(**kw).to(self.device)
        self._finish_init(lr)

    _MODEL = VRPModel
