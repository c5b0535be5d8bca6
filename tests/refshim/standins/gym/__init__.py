"""Minimal stand-in for gym 0.21 (reference mrp00:7-9, setup.py:8): just the names the env modules touch."""
from . import spaces, utils  # noqa: F401


class Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))

    def seed(self, seed=None):
        return [seed]

    def close(self):
        pass
