"""gym_puzzles_b200 — B200-native batched simulator for gym_puzzles' MultiRobotPuzzle family.

Public surface:
  make(id, **kw)                 single env with the reference's gym API (ids of gym_puzzles/__init__.py:3-29)
  VectorEnv(id, num_envs, ...)   device-resident batch exchanging torch CUDA tensors zero-copy
  registry / spec(id)            max_episode_steps, reward_threshold as registered by the reference
"""
from .abi import MrpError, VARIANTS  # noqa: F401
from .envs import MultiRobotPuzzle, MultiRobotPuzzle2, MultiRobotPuzzleHeavy, MultiRobotPuzzleHeavy2  # noqa: F401
from .registry import make, registry, spec  # noqa: F401
from .vector_env import VectorEnv, shard_range  # noqa: F401

__all__ = ["make", "spec", "registry", "VectorEnv", "shard_range", "MultiRobotPuzzle", "MultiRobotPuzzleHeavy",
           "MultiRobotPuzzle2", "MultiRobotPuzzleHeavy2", "MrpError", "VARIANTS"]
