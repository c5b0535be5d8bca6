"""Minimal Box space (gym is optional and not installed in the build image).

Mirrors what the reference declares (reference mrp00:186-207, mrp02:174-195): float32 Boxes; when
`gym` / `gymnasium` is importable the real `spaces.Box` is returned instead so wrappers type-check."""
import numpy as np


class Box:
    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def make_box(low, high, dtype=np.float32):
    for mod in ("gymnasium", "gym"):
        try:
            spaces = __import__(mod + ".spaces", fromlist=["Box"])
            return spaces.Box(np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype), dtype=dtype)
        except Exception:
            continue
    return Box(low, high, dtype)


def observation_space(env_id, n_agents):
    """reference mrp00:186-202 (v0: [inf]*4n + [inf, inf, 2pi, inf] + [inf]*16),
    mrp02:174-190 (v2: [inf, inf, 2pi, inf*6]*n + [inf, inf, 2pi, inf] + [inf]*16 + [inf])."""
    inf, tp = np.inf, 2 * np.pi
    if env_id.endswith("-v0"):
        high = [inf] * 4 * n_agents + [inf, inf, tp, inf] + [inf] * 16
    elif env_id == "MultiRobotPuzzleSquare-v2":   # extension: per block 4 values + its vertices (T 8, L 7, I 4), then 4 scalars
        high = [inf, inf, tp, inf, inf, inf, inf, inf, inf] * n_agents
        for nv in (8, 7, 4):
            high += [inf, inf, tp, inf] + [inf] * (2 * nv)
        high += [inf] * 4
    else:
        high = [inf, inf, tp, inf, inf, inf, inf, inf, inf] * n_agents + [inf, inf, tp, inf] + [inf] * 16 + [inf]
    high = np.array(high, dtype=np.float32)
    return make_box(-high, high)


def action_space(env_id, n_agents):
    """reference mrp00:206-207 (3 per agent), mrp02:194-195 (2 per agent); actions are NOT clipped by the env."""
    k = 3 if env_id.endswith("-v0") else 2
    high = np.ones(k * n_agents, dtype=np.float32)
    return make_box(-high, high)


def reward_defaults(env_id):
    """keyword defaults of set_reward_params (reference mrp00:231-232 / mrp02:216-217): an argument that is not given is
    RESET to its default by the reference, not kept"""
    if env_id.endswith("-v2"):
        return dict(agentDelta=10, agentDistance=0.25, blockDelta=25, blockDistance=0.1, puzzleComp=10000, outOfBounds=1000, blkOutOfBounds=100)
    return dict(agentDelta=10, agentDistance=0.1, blockDelta=50, blockDistance=0.025, puzzleComp=10000, outOfBounds=1000, blkOutOfBounds=100)


REWARD_PARAM_NAMES = ("agentDelta", "agentDistance", "blockDelta", "blockDistance", "puzzleComp", "outOfBounds", "blkOutOfBounds")


def reward_params(env_id, *args, **kwargs):
    """positional / keyword arguments of a set_reward_params call -> the full parameter set the reference would end up with"""
    if len(args) > len(REWARD_PARAM_NAMES):
        raise TypeError("set_reward_params takes at most %d arguments" % len(REWARD_PARAM_NAMES))
    kw = reward_defaults(env_id)
    given = dict(zip(REWARD_PARAM_NAMES, args))
    for k, v in kwargs.items():
        if k not in kw:
            raise TypeError(f"set_reward_params() got an unexpected keyword argument {k!r}")
        if k in given:
            raise TypeError(f"set_reward_params() got multiple values for argument {k!r}")
        given[k] = v
    kw.update(given)
    return kw
