"""VRPEnv — vehicle routing variant: the depot may be revisited (reference gym_vrp/envs/vrp.py:6-37).

The only difference to TSPEnv is mask rule R2 (away from the depot => depot becomes visitable again,
vrp.py:28-31), which the CUDA transition applies when the environment kind is VRP (csrc/env_rules.cuh).
"""
import vrpx

from .tsp import TSPEnv


class VRPEnv(TSPEnv):
    _KIND = vrpx.VRP
