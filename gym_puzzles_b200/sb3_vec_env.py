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


def _cuda_buffers(h, term, n_envs, device):
    """torch views of the library's terminal buffers, a gather of the done rows, pinned host staging buffers"""
    import torch

    from .vector_env import _wrap

    if not h.lib.backend.startswith("cuda"):
        raise abi.MrpError("SB3VecEnv needs the CUDA library: gym_puzzles_b200 has no CPU fallback")
    dev = torch.device("cuda", device)
    t_obs = _wrap(torch, term.terminal_obs_dev, (n_envs, h.obs_dim), "<f4", h, dev)
    t_ret = _wrap(torch, term.episode_return_dev, (n_envs,), "<f4", h, dev)
    t_len = _wrap(torch, term.episode_length_dev, (n_envs,), "<i4", h, dev)

    def gather(idx):
        sel = torch.from_numpy(idx).to(dev)
        return t_obs[sel].cpu().numpy(), t_ret[sel].cpu().numpy(), t_len[sel].cpu().numpy()

    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
    return (t_obs, t_ret, t_len, gather, pin((n_envs, h.obs_dim), torch.float32), pin((n_envs,), torch.float32),
            pin((n_envs,), torch.uint8), pin((n_envs,), torch.uint8), pin((n_envs, h.act_dim), torch.float32))


class SB3VecEnv:
    def __init__(self, env_id, n_envs, device=0, seed=17, n_agents=0, env_id_base=0, max_episode_steps=0, _lib=None, _buffers=None):
        self.handle = abi.Handle(env_id, n_envs, device=device, seed=seed, n_agents=n_agents, env_id_base=env_id_base,
                                 auto_reset=True, max_episode_steps=max_episode_steps, lib=_lib)
        h = self.handle
        self.env_id, self.num_envs = env_id, n_envs
        n = h.layout.n_agents
        self.observation_space = spaces.observation_space(env_id, n)
        self.action_space = spaces.action_space(env_id, n)
        self._term = h.enable_terminal_info()
        # device views of the terminal buffers + pinned host staging buffers; `_buffers` is a test hook (the CPU test-suite
        # injects views over the host build of the kernel source together with `_lib`), the product path is CUDA only
        make = _buffers if _buffers is not None else _cuda_buffers
        (self._t_obs, self._t_ret, self._t_len, self._gather,
         self._obs, self._rew, self._done, self._trunc, self._act) = make(h, self._term, n_envs, device)
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
            t_obs, t_ret, t_len = self._gather(idx)
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
            self.handle.set_params(**spaces.reward_params(self.env_id, *args, **kwargs))
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
