"""Run the UNMODIFIED reference env classes (/root/reference/gym_puzzles/envs/*.py) over the stand-in modules of this
directory.  TEST INFRASTRUCTURE ONLY; used in the build container by tests/test_reference_python.py and
tests/golden/make_golden.py.  Nothing on the GPU box imports this (there is no /root/reference there)."""
import contextlib
import importlib.util
import io
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_STANDINS = os.path.join(_HERE, "standins")   # on sys.path only while a reference module is imported
REFERENCE_ROOT = os.environ.get("MRP_REFERENCE_ROOT", "/root/reference")
STREAM_SPAWN, STREAM_RESET_ACTION = 1, 2          # oracle/philox.hpp
CTOR_EPOCH = 0xFFFFFFFF                            # draws for the reset inside __init__ (mrp00:209): never compared

# registration table of the reference (gym_puzzles/__init__.py:3-29): id -> (module file, class, max_episode_steps)
REGISTRY = {
    "MultiRobotPuzzle-v0": ("multi_robot_puzzle_00.py", "MultiRobotPuzzle", 2000),
    "MultiRobotPuzzleHeavy-v0": ("multi_robot_puzzle_00.py", "MultiRobotPuzzleHeavy", 3000),
    "MultiRobotPuzzle-v2": ("multi_robot_puzzle_02.py", "MultiRobotPuzzle2", 2000),
    "MultiRobotPuzzleHeavy-v2": ("multi_robot_puzzle_02.py", "MultiRobotPuzzleHeavy2", 2000),
}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_puzzles", "envs"))


_modules = {}
_shim = {}    # the stand-in modules; visible in sys.modules only while a reference module is being imported


def _load(fname):
    """Import one reference env module by path with Box2D / gym / pyglet resolving to the stand-ins.  The stand-ins are
    withdrawn from sys.modules afterwards (the reference module keeps its own references), so nothing else in the
    process — in particular gym_puzzles_b200's optional `import gym` — ever sees them."""
    if fname in _modules:
        return _modules[fname]
    hidden = {n: sys.modules.pop(n) for n in list(sys.modules) if n.split(".")[0] in ("Box2D", "gym", "pyglet")}
    sys.modules.update(_shim)
    sys.path.insert(0, _STANDINS)
    try:
        path = os.path.join(REFERENCE_ROOT, "gym_puzzles", "envs", fname)
        spec = importlib.util.spec_from_file_location("reference_" + fname[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(_STANDINS)
        for n in list(sys.modules):
            if n.split(".")[0] in ("Box2D", "gym", "pyglet"):
                _shim[n] = sys.modules.pop(n)
        sys.modules.update(hidden)
    _modules[fname] = mod
    return mod


class ReferenceEnv:
    """One reference env instance + what gym 0.21 wraps around it (TimeLimit from the registration, vector auto-reset),
    with its random draws taken from the Philox streams keyed by (seed, global env id, episode)."""

    def __init__(self, env_id, seed=17, gid=0, max_episode_steps=0, num_agents=None):
        fname, cls, limit = REGISTRY[env_id]
        self.mod = _load(fname)
        self.Box2D, self.gym = _shim["Box2D"], _shim["gym"]
        self.seed, self.gid = seed, gid
        self.max_episode_steps = max_episode_steps or limit
        self.v2 = fname.endswith("02.py")
        self._epoch, self._d = CTOR_EPOCH, 0
        kw = {} if num_agents is None else {"num_agents": num_agents}
        with self._feeds():
            self.env = getattr(self.mod, cls)(**kw)
            if self.v2:
                # the reference leaves shaped_* undefined until update_params is called; decay**(-0) = 1
                self.env.update_params(0, 1.0)
        self.episode = -1
        self.elapsed = 0

    @contextlib.contextmanager
    def _feeds(self):
        u53 = self.Box2D.uniform53

        def uniform(low=0.0, high=1.0, size=None):
            assert size is None
            u = u53(self.seed, STREAM_SPAWN, self.gid, self._epoch, self._d)
            self._d += 1
            return low + (high - low) * u          # numpy: low + (high - low) * random_sample()

        def sampler(shape):
            n = int(np.prod(shape))
            return np.array([np.float32(-1.0 + 2.0 * u53(self.seed, STREAM_RESET_ACTION, self.gid, self._epoch, i)) for i in range(n)],
                            dtype=np.float32).reshape(shape)

        saved, saved_sampler = np.random.uniform, self.gym.spaces.Box.sampler
        np.random.uniform, self.gym.spaces.Box.sampler = uniform, sampler
        try:
            with contextlib.redirect_stdout(io.StringIO()):    # "initialize...", "puzzle complete!!!" (mrp00:195,519)
                yield
        finally:
            np.random.uniform, self.gym.spaces.Box.sampler = saved, saved_sampler

    def reset(self):
        self.episode += 1
        self._epoch, self._d = self.episode, 0
        self.elapsed = 0
        with self._feeds():
            return np.asarray(self.env.reset(), dtype=np.float64)

    def step(self, action):
        """action: float32 values.  They are handed over in a float64 array: NumPy 1.x (gym==0.21 era) evaluates
        np.float32 * python_float in float64; NumPy 2 would not."""
        a = np.asarray(action, dtype=np.float32).astype(np.float64)
        with self._feeds():
            obs, rew, done, info = self.env.step(a)
        self.elapsed += 1
        trunc = False
        if self.elapsed >= self.max_episode_steps:                  # gym.wrappers.TimeLimit
            trunc = not done
            done = True
        obs = np.asarray(obs, dtype=np.float64)
        if done:                                                    # gym 0.21 vector env: reset obs replaces the terminal one
            obs = self.reset()
        return obs, float(rew), bool(done), bool(trunc)

    def set_state(self, layout, words):
        """Put the reference env into the between-steps state of one canonical record (include/mrp_state.h): Box2D side
        through the stand-in's load_state, Python side by assigning the attributes step() reads (goal_contact flags,
        previous distances, goal, blks_in_place).  v0's goal is a module constant and must agree with the record."""
        from oracle_lib import StateView
        sv = StateView(layout, np.asarray(words, dtype=np.uint32).reshape(1, -1))
        w, l, env = sv.w[0], layout, self.env
        nc = int(w[3])
        env.world.load_state(sv.bodies[0], sv.aabbs[0], sv.contacts[0][:nc])
        self.elapsed, self.episode = int(w[0]), int(np.int32(w[1]))
        env.blks_in_place = int(w[2])
        d = sv.dists[0]
        for i, a in enumerate(env.agents):
            a.goal_contact = bool(w[l.off_goal_contact + i])
            env.agent_dist[a.userData] = float(d[i])
        env.block_distance[env.goal_block.userData] = float(d[l.n_agents])
        gx, gy = np.ascontiguousarray(w[l.off_goal:l.off_goal + 4]).view(np.float64)
        if self.v2:
            env.block_final_pos = {env.goal_block.userData: (float(gx), float(gy), 0)}
        else:
            fx, fy, _ = env.block_final_pos[env.goal_block.userData]
            assert (fx, fy) == (gx, gy), "v0 goal is fixed (mrp00:115-128)"

    @property
    def goal_contacts(self):
        return [bool(a.goal_contact) for a in self.env.agents]

    def body_rows(self):
        """(c.x, c.y, a, v.x, v.y, w) float32 of block then agents: the dynamic-body part of include/mrp_state.h"""
        rows = []
        for b in [self.env.goal_block] + list(self.env.agents):
            c, v = b.worldCenter, b.linearVelocity
            rows.append([c[0], c[1], b.angle, v[0], v[1], b.angularVelocity])
        return np.asarray(rows, dtype=np.float32)
