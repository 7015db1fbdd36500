"""ORACLE / TEST INFRASTRUCTURE — minimal stand-in for gym==0.17.2 (only what molgym/spaces.py:21-31,47-53,
77-83,96-101 and molgym/environment.py:17 touch)."""
from . import spaces  # noqa: F401


class Env:
    metadata = {}
    reward_range = (-float('inf'), float('inf'))
    action_space = None
    observation_space = None

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self, mode='human'):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return
