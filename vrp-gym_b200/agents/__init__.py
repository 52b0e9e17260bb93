from .graph_irp_agent import IRPAgent
from .graph_vrp_agent import VRPAgent
from .graph_tsp_agent import TSPAgent
from .random_agent import RandomAgent
from .graph_decoder import GraphDecoder
from .graph_encoder import GraphEncoder, GraphDemandEncoder

__all__ = ["IRPAgent", "VRPAgent", "TSPAgent", "RandomAgent", "GraphDecoder", "GraphEncoder", "GraphDemandEncoder"]
