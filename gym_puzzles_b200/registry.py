"""Env-id registry mirroring reference gym_puzzles/__init__.py:3-29 (gym itself is optional).

`MultiRobotPuzzle-v3` (reference gym_puzzles/__init__.py:31-36, envs/core.py) is outside the hot path named by
BASELINE.json and is not provided."""
from collections import namedtuple

from . import envs

EnvSpec = namedtuple("EnvSpec", "id entry_point max_episode_steps reward_threshold")

registry = {
    "MultiRobotPuzzle-v0": EnvSpec("MultiRobotPuzzle-v0", envs.MultiRobotPuzzle, 2000, 500),
    "MultiRobotPuzzleHeavy-v0": EnvSpec("MultiRobotPuzzleHeavy-v0", envs.MultiRobotPuzzleHeavy, 3000, 500),
    "MultiRobotPuzzle-v2": EnvSpec("MultiRobotPuzzle-v2", envs.MultiRobotPuzzle2, 2000, 500),
    "MultiRobotPuzzleHeavy-v2": EnvSpec("MultiRobotPuzzleHeavy-v2", envs.MultiRobotPuzzleHeavy2, 2000, 500),
    # extension, not in the reference's registry: BASELINE.json configs[4] (three blocks forming a square, Heavy-v2 dynamics)
    "MultiRobotPuzzleSquare-v2": EnvSpec("MultiRobotPuzzleSquare-v2", envs.MultiRobotPuzzleSquare2, 2000, 500),
}


def spec(env_id):
    if env_id not in registry:
        raise KeyError(f"No registered env with id: {env_id}")
    return registry[env_id]


def make(env_id, **kwargs):
    """gym.make equivalent: returns the env with TimeLimit(max_episode_steps) already folded in."""
    s = spec(env_id)
    env = s.entry_point(**kwargs)
    env.spec = s
    return env


def register_with_gym():
    """Optional: expose the same ids through gym / gymnasium when one is installed."""
    done = []
    for mod in ("gymnasium", "gym"):
        try:
            reg = __import__(mod + ".envs.registration", fromlist=["register"]).register
        except Exception:
            continue
        for s in registry.values():
            try:
                reg(id=s.id, entry_point=f"gym_puzzles_b200.envs:{s.entry_point.__name__}", reward_threshold=s.reward_threshold)
                done.append((mod, s.id))
            except Exception:
                pass
    return done
