"""Render-only stub of `gym` so the unmodified reference imports in a container
without gym installed. TEST INFRASTRUCTURE ONLY (oracle/): the reference uses
gym.Env purely as a base class (gym_vrp/envs/tsp.py:4,11)."""


class Env:
    metadata = {}
