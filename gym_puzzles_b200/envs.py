"""Single-environment classes with the reference's gym surface (reset / step / seed / spaces / knobs).

Drop-in for `gym.make('MultiRobotPuzzle-v0')` etc. (reference gym_puzzles/__init__.py:3-29,
gym_puzzles/envs/__init__.py:1-2): each instance is a 1-env C-ABI handle on the GPU; `step` takes / returns numpy
like the reference (obs float32 instead of the reference's float64 array — SURVEY.md C.8).  TimeLimit is folded in
(`info['TimeLimit.truncated']`).  Rendering is out of scope (host-side viewer code stays with the reference).
For throughput use `VectorEnv`; this class pays one kernel launch + two PCIe copies per step by construction.
"""
import numpy as np

from . import abi, spaces


class MultiRobotPuzzle:
    env_id = "MultiRobotPuzzle-v0"
    metadata = {"render.modes": [], "video.frames_per_second": 50}
    reward_range = (-float("inf"), float("inf"))
    spec = None

    def __init__(self, obs_depth=3, frameskip=4, num_agents=0, seed=17, device=0, _lib=None):
        # obs_depth / frameskip are accepted for signature compatibility: the v0 classes force frameskip to 1 for the
        # low-dim observation (reference mrp00:161-162), whatever is passed
        self._seed_value = seed
        self._device = device
        self._lib = _lib
        self._num_agents = num_agents
        self._episode_base = 0
        self._reward_kw, self._knob_kw = None, {}   # knob state survives seed() (which rebuilds the handle)
        self._make_handle()
        self.observation_space = spaces.observation_space(self.env_id, self.num_agents)
        self.action_space = spaces.action_space(self.env_id, self.num_agents)
        self.done_status = None
        self._last_done = False

    def _make_handle(self):
        self._h = abi.Handle(self.env_id, 1, device=self._device, seed=self._seed_value, n_agents=self._num_agents,
                             auto_reset=False, lib=self._lib)
        self.num_agents = self._h.layout.n_agents
        if self._reward_kw:
            self._h.set_params(**self._reward_kw)
        if self._knob_kw:
            self._h.set_params(**self._knob_kw)

    # ---- gym API
    def seed(self, seed=None):
        """reference mrp00:211-216.  Here the seed keys the Philox spawn stream, so (unlike the reference, whose
        spawns use numpy's global RNG — SURVEY.md C.2) it does make resets reproducible."""
        if seed is None:
            seed = int(np.random.SeedSequence().entropy % (2 ** 63))
        self._seed_value = int(seed)
        self._h.close()
        self._make_handle()
        return [self._seed_value]

    def reset(self):
        self.done_status = None
        self._last_done = False
        return self._h.reset_host()[0].copy()

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, -1)
        if a.shape[1] != self._h.act_dim:
            raise ValueError(f"action must have {self._h.act_dim} entries, got {a.shape[1]}")
        obs, rew, done, trunc = self._h.step_host(a)
        info = {}
        if done[0]:
            if trunc[0]:
                info["TimeLimit.truncated"] = True
            else:
                self.done_status = "puzzle complete!!" if self.env_id.endswith("-v0") else "episode finished"
        self._last_done = bool(done[0])
        return obs[0].copy(), float(rew[0]), bool(done[0]), info

    def render(self, mode="human", close=False):
        """reference mrp00:528-592 / mrp02:590-707.  'rgb_array' is drawn on the host from get_state() by
        gym_puzzles_b200.render (numpy, no pyglet); an interactive window ('human') is not provided."""
        if close:
            return None
        if mode == "rgb_array":
            from . import render as _render
            return _render.rgb_array(self._h, 0)
        raise NotImplementedError("only render(mode='rgb_array') is provided; rendering is host-side and outside the B200 hot path")

    def close(self):
        self._h.close()

    # ---- knobs (reference mrp00:231-258 / mrp02:216-245)
    def set_reward_params(self, *args, **kwargs):
        """reference mrp00:231-239 / mrp02:216-225; arguments not given return to the family's defaults, as there"""
        self._reward_kw = spaces.reward_params(self.env_id, *args, **kwargs)
        self._h.set_params(**self._reward_kw)

    def update_params(self, timestep, decay):
        self._knob_kw["decay_pow"] = float(decay) ** (-float(timestep))
        self._h.set_params(decay_pow=self._knob_kw["decay_pow"])

    def update_goal(self, epoch, nb_epochs):
        self._knob_kw["scaled_epsilon"] = 0.1 * (2 - epoch / nb_epochs)
        self._h.set_params(scaled_epsilon=self._knob_kw["scaled_epsilon"])

    def get_deltaAgent(self):
        return self._h.get_params()["agentDelta"]

    def get_agentDist(self):
        return self._h.get_params()["agentDistance"]

    def get_deltaBlk(self):
        return self._h.get_params()["blockDelta"]

    def get_blkDist(self):
        return self._h.get_params()["blockDistance"]

    def _return_status(self):
        return self.done_status if self.done_status else "Stayed in bounds"

    # ---- state (new: the reference cannot serialise an env)
    def get_state(self):
        return self._h.get_state()[0]

    def set_state(self, words):
        self._h.set_state(np.asarray(words, dtype=np.uint32).reshape(1, -1))

    @property
    def unwrapped(self):
        return self


class MultiRobotPuzzleHeavy(MultiRobotPuzzle):
    env_id = "MultiRobotPuzzleHeavy-v0"   # heavy = True, number_agents = 5 (reference mrp00:606-610)


class MultiRobotPuzzle2(MultiRobotPuzzle):
    env_id = "MultiRobotPuzzle-v2"

    def __init__(self, frameskip=1, num_agents=2, seed=17, device=0, _lib=None):
        # reference mrp02:477-478 runs world.Step `frameskip` times per env.step; the kernels run it once
        if frameskip != 1:
            raise NotImplementedError("MultiRobotPuzzle2(frameskip != 1) is not provided: one world.Step per env.step (the registered default)")
        super().__init__(frameskip=frameskip, num_agents=num_agents, seed=seed, device=device, _lib=_lib)


class MultiRobotPuzzleHeavy2(MultiRobotPuzzle2):
    env_id = "MultiRobotPuzzleHeavy-v2"   # heavy = True: block density 20 (reference mrp02:162-163,711-712)


class MultiRobotPuzzleSquare2(MultiRobotPuzzle2):
    """Extension (BASELINE.json configs[4]), NOT a reference env: the T, L and I blocks of reference mrp00:320-351 /
    blocks.py:70-109 with Heavy-v2 dynamics; the blocks are pushed, in the reference's block_queue order, to the poses of
    mrp00:83-88 that tile a square around the goal (semantics: DESIGN.md "Square variant")."""
    env_id = "MultiRobotPuzzleSquare-v2"
