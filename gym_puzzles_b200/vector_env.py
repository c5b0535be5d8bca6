"""VectorEnv — the batched, device-resident face of the MultiRobotPuzzle family.

`num_envs` environments live in HBM inside one C-ABI handle (include/mrp_b200.h); `reset()` / `step()` exchange
torch CUDA tensors that alias the library's buffers (zero-copy).  Semantics per env follow the reference
`reset()` / `step()` (reference mrp00:392-521, mrp02:421-584) wrapped in gym's TimeLimit
(gym_puzzles/__init__.py:3-29) with gym-0.21 vector auto-reset: when an env is done its row of `obs` already
holds the first observation of the next episode.
"""
import os

import numpy as np

from . import abi, spaces


class _DevArray:
    """__cuda_array_interface__ view of a library-owned device buffer (kept alive by the handle)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._owner = owner


def _wrap(torch, ptr, shape, typestr, owner, device):
    return torch.as_tensor(_DevArray(ptr, shape, typestr, owner), device=device)


class VectorEnv:
    """VectorEnv(env_id, num_envs, device='cuda:0', seed=17)

    reset() -> obs[N, O]                                   (torch.float32, CUDA, aliases the library buffer)
    step(actions[N, A]) -> obs[N, O], reward[N], done[N] (bool), info {'TimeLimit.truncated': trunc[N] (bool)}
    """

    def __init__(self, env_id, num_envs, device="cuda:0", seed=17, n_agents=0, env_id_base=0, auto_reset=True,
                 max_episode_steps=0):
        import torch

        self.torch = torch
        self.env_id = env_id
        if not torch.cuda.is_available():
            raise abi.MrpError("VectorEnv needs a CUDA device: gym_puzzles_b200 has no CPU fallback")
        self.device = torch.device(device)
        ordinal = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.handle = abi.Handle(env_id, num_envs, device=ordinal, seed=seed, n_agents=n_agents, env_id_base=env_id_base,
                                 auto_reset=auto_reset, max_episode_steps=max_episode_steps)
        h, b = self.handle, self.handle.buffers
        self.num_envs = num_envs
        self.n_agents = h.layout.n_agents
        self.single_observation_space = spaces.observation_space(env_id, self.n_agents)
        self.single_action_space = spaces.action_space(env_id, self.n_agents)
        dev = self.device
        self.actions = _wrap(torch, b.action_dev, (num_envs, h.act_dim), "<f4", h, dev)
        self.obs = _wrap(torch, b.obs_dev, (num_envs, h.obs_dim), "<f4", h, dev)
        self.reward = _wrap(torch, b.reward_dev, (num_envs,), "<f4", h, dev)
        # the library writes 0/1 bytes: exposed as torch.bool without a conversion kernel
        self.done = _wrap(torch, b.done_dev, (num_envs,), "|b1", h, dev)
        self.truncated = _wrap(torch, b.trunc_dev, (num_envs,), "|b1", h, dev)
        self.stats_tensor = _wrap(torch, b.stats_dev, (abi.N_STATS,), "<f8", h, dev)
        self._step_index = 0
        self._stats_step0 = 0
        self.terminal_obs = self.episode_return = self.episode_length = None
        self.scaled_epsilon = self.decay_pow = None

    # ------------------------------------------------------------------ SURVEY.md §8f "next" rows
    def enable_terminal_info(self):
        """Keep, for envs that finish inside step(), the last observation of the episode and its return / length
        (SB3's infos[i]["terminal_observation"], Monitor's info["episode"]): `terminal_obs[N, O]`,
        `episode_return[N]`, `episode_length[N]` — rows are valid where `done` is set by the same step."""
        if self.terminal_obs is None:
            t, h, dev, torch = self.handle.enable_terminal_info(), self.handle, self.device, self.torch
            self.terminal_obs = _wrap(torch, t.terminal_obs_dev, (self.num_envs, h.obs_dim), "<f4", h, dev)
            self.episode_return = _wrap(torch, t.episode_return_dev, (self.num_envs,), "<f4", h, dev)
            self.episode_length = _wrap(torch, t.episode_length_dev, (self.num_envs,), "<i4", h, dev)
        return self.terminal_obs, self.episode_return, self.episode_length

    def enable_curriculum(self):
        """Per-env curriculum vectors (float64 CUDA tensors [N], written by the caller, read by the next step):
        `scaled_epsilon` = update_goal's tolerance (mrp02:232-233), `decay_pow` = decay**(-timestep) of update_params
        (mrp02:227-230).  After this, update_goal / update_params also accept per-env arrays."""
        if self.scaled_epsilon is None:
            e, d = self.handle.enable_curriculum()
            self.scaled_epsilon = _wrap(self.torch, e, (self.num_envs,), "<f8", self.handle, self.device)
            self.decay_pow = _wrap(self.torch, d, (self.num_envs,), "<f8", self.handle, self.device)
        return self.scaled_epsilon, self.decay_pow

    # ------------------------------------------------------------------ core API
    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, mask=None):
        """Respawn (all envs, or those where mask[N] is nonzero) and return obs.  Stream-ordered, no host sync."""
        mptr = None
        if mask is not None:
            self._mask = mask.to(device=self.device, dtype=self.torch.uint8).contiguous()
            mptr = self._mask.data_ptr()
        self.handle.reset(mptr, self._stream())
        return self.obs

    def step(self, actions=None):
        """actions: float32 CUDA tensor [N, A] (any tensor is read in place, no copy); None = the owned buffer
        `self.actions` (write into it to avoid even the pointer hand-off)."""
        aptr = None
        if actions is not None:
            if actions.device != self.device or actions.dtype != self.torch.float32 or not actions.is_contiguous():
                actions = actions.to(device=self.device, dtype=self.torch.float32).contiguous()
            if tuple(actions.shape) != (self.num_envs, self.handle.act_dim):
                raise ValueError(f"actions must have shape {(self.num_envs, self.handle.act_dim)}, got {tuple(actions.shape)}")
            self._last_actions = actions  # keep alive until the kernel has run
            aptr = actions.data_ptr()
        self.handle.step(aptr, self._stream())
        self._step_index += 1
        return self.obs, self.reward, self.done, {"TimeLimit.truncated": self.truncated}

    def obs_v3(self, out=None):
        """The normalised observation of the reference's experimental MultiRobotPuzzle-v3 (gym_puzzles/envs/core.py:289-350)
        for the current state of a v0 / Heavy-v0 batch: float32 CUDA tensor [N, 4 n + 19]."""
        if out is None:
            out = self.torch.empty((self.num_envs, 4 * self.n_agents + 19), dtype=self.torch.float32, device=self.device)
        self.handle.obs_v3(out.data_ptr(), self._stream())
        return out

    def sample_actions(self, step_index=None, out=None):
        """Synthetic U(-1,1) actions (Philox stream ACTION keyed by global env id and step) written on the device."""
        idx = self._step_index if step_index is None else step_index
        self.handle.sample_actions(idx, None if out is None else out.data_ptr(), self._stream())
        return self.actions if out is None else out

    def close(self):
        self.handle.close()

    # ------------------------------------------------------------------ host-buffer path (numpy in / numpy out)
    def step_host(self, actions, **out):
        return self.handle.step_host(actions, **out)

    def reset_host(self, mask=None):
        return self.handle.reset_host(mask)

    # ------------------------------------------------------------------ knobs of the reference env classes
    def set_reward_params(self, *args, **kwargs):
        """reference mrp00:231-239 / mrp02:216-225; arguments not given return to the family's defaults, as there"""
        self.handle.set_params(**spaces.reward_params(self.env_id, *args, **kwargs))

    def _per_env(self, x):
        t = self.torch
        return t.is_tensor(x) or (hasattr(x, "__len__") and len(x) == self.num_envs)

    def update_params(self, timestep, decay):
        """reference mrp02:227-230: shaped rewards = base * decay**(-timestep).  Arrays / tensors of length N set
        the value per env (needs enable_curriculum(); computed in float64 like the reference's Python)."""
        if self._per_env(timestep) or self._per_env(decay):
            self.enable_curriculum()
            t = self.torch
            ts = t.as_tensor(timestep, dtype=t.float64, device=self.device)
            dc = t.as_tensor(decay, dtype=t.float64, device=self.device)
            self.decay_pow.copy_((dc ** (-ts)).expand(self.num_envs))
            return
        v = float(decay) ** (-float(timestep))
        self.handle.set_params(decay_pow=v)
        if self.decay_pow is not None:
            self.decay_pow.fill_(v)

    def update_goal(self, epoch, nb_epochs):
        """reference mrp02:232-233: scaled_epsilon = EPSILON * (2 - epoch / nb_epochs), EPSILON = 0.1."""
        if self._per_env(epoch):
            self.enable_curriculum()
            t = self.torch
            ep = t.as_tensor(epoch, dtype=t.float64, device=self.device)
            self.scaled_epsilon.copy_(0.1 * (2 - ep / nb_epochs))
            return
        v = 0.1 * (2 - epoch / nb_epochs)
        self.handle.set_params(scaled_epsilon=v)
        if self.scaled_epsilon is not None:
            self.scaled_epsilon.fill_(v)

    def get_deltaAgent(self):
        return self.handle.get_params()["agentDelta"]

    def get_agentDist(self):
        return self.handle.get_params()["agentDistance"]

    def get_deltaBlk(self):
        return self.handle.get_params()["blockDelta"]

    def get_blkDist(self):
        return self.handle.get_params()["blockDistance"]

    # ------------------------------------------------------------------ state & statistics
    def get_state(self, begin=0, count=None):
        return self.handle.get_state(begin, count)

    def set_state(self, words, begin=0):
        self.handle.set_state(words, begin)

    def episode_stats(self, reduce_across_ranks=True, reset=True):
        """Episode statistics accumulated on the device; summed over ranks with one NCCL all-reduce of 16 doubles
        when torch.distributed is initialised (the only collective on this path — SURVEY.md §8e)."""
        out = reduce_stats(self.torch, self.stats_tensor, reduce_across_ranks, reset,
                           env_steps=(self._step_index - self._stats_step0) * self.num_envs)
        if reset:
            self._stats_step0 = self._step_index
        return out


def reduce_stats(torch, stats_tensor, reduce_across_ranks=True, reset=True, env_steps=0):
    s = stats_tensor.clone()
    s[abi.STAT_NAMES.index("env_steps")] = float(env_steps)   # counted on the host (no per-env atomics on the device)
    if reset:
        stats_tensor.zero_()
    if reduce_across_ranks:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
    v = s.tolist()
    out = dict(zip(abi.STAT_NAMES, v))
    n = max(out["episodes"], 1.0)
    out["mean_return"] = out["sum_return"] / n
    out["var_return"] = max(out["sum_return_sq"] / n - out["mean_return"] ** 2, 0.0)
    out["mean_length"] = out["sum_length"] / n
    return out


def shard_range(total_envs, rank, world_size):
    """Contiguous env-id block of `rank`: env g lives on rank g // (total / world)."""
    if total_envs % world_size:
        raise ValueError("total_envs must be divisible by world_size")
    per = total_envs // world_size
    return rank * per, per
