"""SB3VecEnv — the batched simulator behind Stable-Baselines3's `VecEnv` calling convention, i.e. what the
reference's trainer builds with `DummyVecEnv([make_env] * n_envs)` over `Monitor(gym.make(id))`
(reference train/train.py:63-82).  numpy in, numpy out, one `mrp_step_host` call per `step_wait`:

    reset() -> obs[n, O]
    step_async(actions[n, A]); step_wait() -> obs[n, O], rewards[n], dones[n] (bool), infos (list of n dicts)
    step(actions)                           = step_async + step_wait

For an env that finished inside the step, `obs[i]` already is the first observation of its next episode (SB3's
vector-env contract) and `infos[i]` carries
    "terminal_observation"   last observation of the finished episode,
    "TimeLimit.truncated"    True when the registered max_episode_steps ended it (gym TimeLimit),
    "episode": {"r", "l", "t"}   return / length / wall time since construction (SB3 Monitor, train/train.py:68).
SB3 itself is not imported (it is not a dependency of this package); the class is duck-type compatible and registers
as a virtual subclass of `stable_baselines3.common.vec_env.VecEnv` when SB3 is importable."""
import time

import numpy as np

from . import abi, spaces


class SB3VecEnv:
    def __init__(self, env_id, n_envs, device=0, seed=17, n_agents=0, env_id_base=0, max_episode_steps=0, _lib=None):
        self.handle = abi.Handle(env_id, n_envs, device=device, seed=seed, n_agents=n_agents, env_id_base=env_id_base,
                                 auto_reset=True, max_episode_steps=max_episode_steps, lib=_lib)
        h = self.handle
        self.env_id, self.num_envs = env_id, n_envs
        n = h.layout.n_agents
        self.observation_space = spaces.observation_space(env_id, n)
        self.action_space = spaces.action_space(env_id, n)
        self._term = h.enable_terminal_info()
        self._cuda = h.lib.backend.startswith("cuda")
        if self._cuda:
            import torch
            from .vector_env import _wrap

            dev = torch.device("cuda", device)
            self._t_obs = _wrap(torch, self._term.terminal_obs_dev, (n_envs, h.obs_dim), "<f4", h, dev)
            self._t_ret = _wrap(torch, self._term.episode_return_dev, (n_envs,), "<f4", h, dev)
            self._t_len = _wrap(torch, self._term.episode_length_dev, (n_envs,), "<i4", h, dev)
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
            self._obs, self._rew = pin((n_envs, h.obs_dim), torch.float32), pin((n_envs,), torch.float32)
            self._done, self._trunc = pin((n_envs,), torch.uint8), pin((n_envs,), torch.uint8)
            self._act = pin((n_envs, h.act_dim), torch.float32)
        else:   # host build of the kernel source (tests only): "device" pointers are host pointers
            import ctypes as C

            view = lambda p, shape, ct, dt: np.frombuffer((ct * int(np.prod(shape))).from_address(p), dtype=dt).reshape(shape)  # noqa: E731
            self._t_obs = view(self._term.terminal_obs_dev, (n_envs, h.obs_dim), C.c_float, np.float32)
            self._t_ret = view(self._term.episode_return_dev, (n_envs,), C.c_float, np.float32)
            self._t_len = view(self._term.episode_length_dev, (n_envs,), C.c_int32, np.int32)
            self._obs, self._rew = np.empty((n_envs, h.obs_dim), np.float32), np.empty(n_envs, np.float32)
            self._done, self._trunc = np.empty(n_envs, np.uint8), np.empty(n_envs, np.uint8)
            self._act = np.empty((n_envs, h.act_dim), np.float32)
        self._t0 = time.time()
        self._pending = False
        self.render_mode = None

    # ---- VecEnv API
    def reset(self):
        return self.handle.reset_host(None, self._obs).copy()

    def step_async(self, actions):
        self._act[...] = np.asarray(actions, dtype=np.float32).reshape(self._act.shape)
        self._pending = True

    def step_wait(self):
        assert self._pending, "step_wait() without step_async()"
        self._pending = False
        self.handle.step_host(self._act, self._obs, self._rew, self._done, self._trunc)
        dones = self._done.astype(bool)
        infos = [{} for _ in range(self.num_envs)]
        idx = np.nonzero(dones)[0]
        if idx.size:
            if self._cuda:
                import torch

                sel = torch.from_numpy(idx).to(self._t_obs.device)
                t_obs, t_ret, t_len = self._t_obs[sel].cpu().numpy(), self._t_ret[sel].cpu().numpy(), self._t_len[sel].cpu().numpy()
            else:
                t_obs, t_ret, t_len = self._t_obs[idx].copy(), self._t_ret[idx].copy(), self._t_len[idx].copy()
            now = round(time.time() - self._t0, 6)
            for k, i in enumerate(idx):
                infos[i] = {"terminal_observation": t_obs[k], "TimeLimit.truncated": bool(self._trunc[i]),
                            "episode": {"r": float(t_ret[k]), "l": int(t_len[k]), "t": now}}
        return self._obs.copy(), self._rew.copy(), dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.handle.close()

    def seed(self, seed=None):
        """Spawns are drawn from Philox keyed by (seed, global env id, episode); as in the reference, env.seed() does not
        re-seed them (the reference spawns from the global np.random)."""
        return [seed] * self.num_envs

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    # the reference's runtime knobs, addressed the way SB3 does it: venv.env_method("set_reward_params", ...)
    def env_method(self, method_name, *args, indices=None, **kwargs):
        if method_name == "set_reward_params":
            names = ("agentDelta", "agentDistance", "blockDelta", "blockDistance", "puzzleComp", "outOfBounds", "blkOutOfBounds")
            kw = dict(zip(names, args))
            kw.update(kwargs)
            self.handle.set_params(**kw)
        elif method_name == "update_params":
            timestep, decay = args
            self.handle.set_params(decay_pow=float(decay) ** (-float(timestep)))
        elif method_name == "update_goal":
            epoch, nb_epochs = args
            self.handle.set_params(scaled_epsilon=0.1 * (2 - epoch / nb_epochs))
        else:
            raise AttributeError(f"env_method({method_name!r}) is not provided by the batched env")
        return [None] * self.num_envs

    def get_attr(self, attr_name, indices=None):
        p = self.handle.get_params()
        table = {"weight_deltaAgent": "agentDelta", "weight_agent_dist": "agentDistance", "weight_deltaBlock": "blockDelta",
                 "weight_blk_dist": "blockDistance", "puzzle_complete_reward": "puzzleComp", "scaled_epsilon": "scaled_epsilon"}
        if attr_name in table:
            return [p[table[attr_name]]] * self.num_envs
        if attr_name == "render_mode":
            return [None] * self.num_envs
        raise AttributeError(attr_name)

    def set_attr(self, attr_name, value, indices=None):
        raise AttributeError("use env_method('set_reward_params' | 'update_params' | 'update_goal', ...)")

    def get_images(self):
        raise NotImplementedError("rendering is host-side and per env: gym_puzzles_b200.render.scene(handle, i)")

    def render(self, mode="human"):
        raise NotImplementedError("rendering is host-side and per env: gym_puzzles_b200.render.rgb_array(handle, i)")


try:   # optional: isinstance(venv, VecEnv) for SB3 code paths that check it
    from stable_baselines3.common.vec_env import VecEnv as _VecEnv  # type: ignore

    _VecEnv.register(SB3VecEnv)
except Exception:
    pass
