"""gym_puzzles_b200 — B200-native batched simulator for gym_puzzles' MultiRobotPuzzle family.

Public surface:
  make(id, **kw)                 single env with the reference's gym API (ids of gym_puzzles/__init__.py:3-29)
  VectorEnv(id, num_envs, ...)   device-resident batch exchanging torch CUDA tensors zero-copy
  registry / spec(id)            max_episode_steps, reward_threshold as registered by the reference
  SB3VecEnv(id, n_envs)          Stable-Baselines3 VecEnv calling convention (numpy, terminal_observation, Monitor info)
  VecNormalize(venv)             device-resident running obs / return normalisation (sm_100a kernels)
  render.scene / rgb_array       host render bridge from get_state()
"""
from .abi import MrpError, VARIANTS  # noqa: F401
from .envs import MultiRobotPuzzle, MultiRobotPuzzle2, MultiRobotPuzzleHeavy, MultiRobotPuzzleHeavy2, MultiRobotPuzzleSquare2  # noqa: F401
from .registry import make, register_with_gym, registry, spec  # noqa: F401
from .vector_env import VectorEnv, shard_range  # noqa: F401
from .sb3_vec_env import SB3VecEnv  # noqa: F401
from .vec_normalize import VecNormalize  # noqa: F401
from . import render  # noqa: F401

# the reference registers its ids with gym on import (gym_puzzles/__init__.py:3-29): do the same where gym / gymnasium
# is installed, so that gym.make("MultiRobotPuzzleHeavy-v0") resolves to this package after `import gym_puzzles_b200`
GYM_REGISTERED = register_with_gym()

__all__ = ["make", "spec", "registry", "register_with_gym", "VectorEnv", "shard_range", "MultiRobotPuzzle", "MultiRobotPuzzleHeavy",
           "MultiRobotPuzzle2", "MultiRobotPuzzleHeavy2", "MultiRobotPuzzleSquare2", "MrpError", "VARIANTS", "SB3VecEnv", "VecNormalize", "render"]
