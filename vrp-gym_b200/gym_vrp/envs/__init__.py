from .irp import IRPEnv
from .tsp import TSPEnv
from .vrp import VRPEnv

__all__ = ["TSPEnv", "VRPEnv", "IRPEnv"]
