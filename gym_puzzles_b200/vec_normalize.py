"""Device-resident VecNormalize for VectorEnv — what the reference's trainer wraps around its envs
(reference train/train.py:82 `env = VecNormalize(env)`, reloaded for evaluation at train/test.py:66-68), computed by
the sm_100a kernels of csrc/mrp_vecnorm.cu on the obs / reward tensors where they already live (HBM).

Semantics follow Stable-Baselines3's VecNormalize (a third-party dependency of the reference's training script):
running mean / variance of observations and of discounted returns, `clip((obs - mean) / sqrt(var + eps), ±clip_obs)`,
`clip(reward / sqrt(var_ret + eps), ±clip_reward)`, statistics updated before normalising while `training`, returns
zeroed where done.  With torch.distributed initialised the batch moments are summed over ranks with ONE NCCL
all-reduce per step, so every rank keeps identical statistics (as one VecNormalize over the global batch would)."""
import ctypes as C

import numpy as np

from . import abi


class VnConfig(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("obs_dim", C.c_int32), ("device", C.c_int32), ("norm_obs", C.c_int32),
                ("norm_reward", C.c_int32), ("training", C.c_int32), ("clip_obs", C.c_double), ("clip_reward", C.c_double),
                ("gamma", C.c_double), ("epsilon", C.c_double)]


VN_EXPORTS = ("mrp_vecnorm_last_error", "mrp_vecnorm_create", "mrp_vecnorm_destroy", "mrp_vecnorm_set_training",
              "mrp_vecnorm_moments", "mrp_vecnorm_accum", "mrp_vecnorm_apply", "mrp_vecnorm_reset_returns",
              "mrp_vecnorm_get_stats", "mrp_vecnorm_set_stats", "mrp_vecnorm_launch_count")


def _bind(lib):
    L = lib.lib
    if getattr(L, "_vn_bound", False):
        return L
    p = C.c_void_p
    L.mrp_vecnorm_last_error.restype = C.c_char_p
    L.mrp_vecnorm_create.argtypes = [C.POINTER(VnConfig), C.POINTER(p)]
    L.mrp_vecnorm_destroy.argtypes = [p]
    L.mrp_vecnorm_set_training.argtypes = [p, C.c_int32]
    L.mrp_vecnorm_moments.argtypes = [p, p, p, p]
    L.mrp_vecnorm_accum.argtypes = [p, C.POINTER(p), C.POINTER(C.c_int32)]
    L.mrp_vecnorm_apply.argtypes = [p, p, p, p, p, p, p, p]
    L.mrp_vecnorm_reset_returns.argtypes = [p, p]
    L.mrp_vecnorm_get_stats.argtypes = [p, p]
    L.mrp_vecnorm_set_stats.argtypes = [p, p]
    L.mrp_vecnorm_launch_count.argtypes = [p]
    L.mrp_vecnorm_launch_count.restype = C.c_int64
    L._vn_bound = True
    return L


class VecNormHandle:
    """Thin owner of one mrp_vecnorm (pointers in, pointers out; used directly by tests through the host emulation)."""

    def __init__(self, num_envs, obs_dim, device=0, norm_obs=True, norm_reward=True, training=True, clip_obs=10.0,
                 clip_reward=10.0, gamma=0.99, epsilon=1e-8, lib=None):
        self.lib = lib if lib is not None else abi.load()
        self.L = _bind(self.lib)
        cfg = VnConfig(num_envs, obs_dim, device, int(norm_obs), int(norm_reward), int(training), clip_obs, clip_reward, gamma, epsilon)
        self.h = C.c_void_p()
        self._check(self.L.mrp_vecnorm_create(C.byref(cfg), C.byref(self.h)), "mrp_vecnorm_create")
        self.num_envs, self.obs_dim = num_envs, obs_dim

    def _check(self, rc, what):
        if rc != 0:
            raise abi.MrpError(f"{what} failed ({rc}): {self.L.mrp_vecnorm_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.mrp_vecnorm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_training(self, on):
        self._check(self.L.mrp_vecnorm_set_training(self.h, int(on)), "mrp_vecnorm_set_training")

    def moments(self, obs_ptr, reward_ptr=None, stream=None):
        self._check(self.L.mrp_vecnorm_moments(self.h, obs_ptr, reward_ptr, stream), "mrp_vecnorm_moments")

    def accum(self):
        p, n = C.c_void_p(), C.c_int32()
        self._check(self.L.mrp_vecnorm_accum(self.h, C.byref(p), C.byref(n)), "mrp_vecnorm_accum")
        return p.value, n.value

    def apply(self, obs_ptr, reward_ptr=None, done_ptr=None, obs_out_ptr=None, reward_out_ptr=None, term_obs_ptr=None, stream=None):
        self._check(self.L.mrp_vecnorm_apply(self.h, obs_ptr, reward_ptr, done_ptr, obs_out_ptr, reward_out_ptr, term_obs_ptr, stream),
                    "mrp_vecnorm_apply")

    def reset_returns(self, stream=None):
        self._check(self.L.mrp_vecnorm_reset_returns(self.h, stream), "mrp_vecnorm_reset_returns")

    def get_stats(self):
        s = np.zeros(2 * (self.obs_dim + 1) + 2, dtype=np.float64)
        self._check(self.L.mrp_vecnorm_get_stats(self.h, s.ctypes.data), "mrp_vecnorm_get_stats")
        return s

    def set_stats(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.shape == (2 * (self.obs_dim + 1) + 2,)
        self._check(self.L.mrp_vecnorm_set_stats(self.h, s.ctypes.data), "mrp_vecnorm_set_stats")

    @property
    def launch_count(self):
        return int(self.L.mrp_vecnorm_launch_count(self.h))


class VecNormalize:
    """VecNormalize(venv) over a VectorEnv: same surface as SB3's wrapper for the calls the reference makes
    (`reset`, `step`, `training`, `norm_reward`, `save` / `load`, `normalize_obs`, `get_original_obs/reward`)."""

    def __init__(self, venv, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0, gamma=0.99,
                 epsilon=1e-8, sync_across_ranks=True):
        torch = venv.torch
        self.venv, self.torch = venv, torch
        self.num_envs = venv.num_envs
        O = venv.handle.obs_dim
        ordinal = venv.device.index if venv.device.index is not None else torch.cuda.current_device()
        self.vn = VecNormHandle(venv.num_envs, O, ordinal, norm_obs, norm_reward, training, clip_obs, clip_reward, gamma, epsilon)
        self._training, self.norm_obs, self.norm_reward = bool(training), bool(norm_obs), bool(norm_reward)
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon
        self.sync_across_ranks = sync_across_ranks
        self.obs = torch.empty((venv.num_envs, O), dtype=torch.float32, device=venv.device)
        self.reward = torch.empty((venv.num_envs,), dtype=torch.float32, device=venv.device)
        from .vector_env import _wrap
        p, n = self.vn.accum()
        self._accum = _wrap(torch, p, (n,), "<f8", self.vn, venv.device)
        self.single_observation_space, self.single_action_space = venv.single_observation_space, venv.single_action_space

    # SB3 attribute: VecNormalize.training (train/test.py:67)
    @property
    def training(self):
        return self._training

    @training.setter
    def training(self, on):
        self._training = bool(on)
        self.vn.set_training(on)

    def _stream(self):
        return self.torch.cuda.current_stream(self.venv.device).cuda_stream

    def _sync_moments(self):
        if self._training and self.sync_across_ranks:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(self._accum, op=dist.ReduceOp.SUM)

    def reset(self):
        obs = self.venv.reset()
        st = self._stream()
        self.vn.reset_returns(st)
        self.vn.moments(obs.data_ptr(), None, st)
        self._sync_moments()
        self.vn.apply(obs.data_ptr(), None, None, self.obs.data_ptr(), None, None, st)
        return self.obs

    def step(self, actions=None):
        obs, rew, done, info = self.venv.step(actions)
        st = self._stream()
        self.vn.moments(obs.data_ptr(), rew.data_ptr(), st)
        self._sync_moments()
        term = getattr(self.venv, "terminal_obs", None)
        self.vn.apply(obs.data_ptr(), rew.data_ptr(), done.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(),
                      None if term is None else term.data_ptr(), st)
        return self.obs, self.reward, done, info

    def get_original_obs(self):
        return self.venv.obs

    def get_original_reward(self):
        return self.venv.reward

    def normalize_obs(self, obs):
        """Normalise an arbitrary obs tensor with the current statistics (no update)."""
        s = self.vn.get_stats()
        O = self.vn.obs_dim
        mean = self.torch.as_tensor(s[:O], dtype=self.torch.float32, device=obs.device)
        istd = self.torch.as_tensor(1.0 / np.sqrt(s[O + 1:2 * O + 1] + self.epsilon), dtype=self.torch.float32, device=obs.device)
        return ((obs - mean) * istd).clamp(-self.clip_obs, self.clip_obs) if self.norm_obs else obs

    # ---- save / load: the numbers SB3 pickles into saved_env.pkl (train/train.py:149), as a portable .npz
    def state_dict(self):
        s = self.vn.get_stats()
        O = self.vn.obs_dim
        return {"obs_rms.mean": s[:O].copy(), "obs_rms.var": s[O + 1:2 * O + 1].copy(), "obs_rms.count": float(s[2 * O + 2]),
                "ret_rms.mean": float(s[O]), "ret_rms.var": float(s[2 * O + 1]), "ret_rms.count": float(s[2 * O + 3]),
                "clip_obs": self.clip_obs, "clip_reward": self.clip_reward, "gamma": self.gamma, "epsilon": self.epsilon,
                "norm_obs": self.norm_obs, "norm_reward": self.norm_reward}

    def load_state_dict(self, d):
        O = self.vn.obs_dim
        s = np.zeros(2 * (O + 1) + 2)
        s[:O], s[O] = d["obs_rms.mean"], d["ret_rms.mean"]
        s[O + 1:2 * O + 1], s[2 * O + 1] = d["obs_rms.var"], d["ret_rms.var"]
        s[2 * O + 2], s[2 * O + 3] = d["obs_rms.count"], d["ret_rms.count"]
        self.vn.set_stats(s)

    def save(self, path):
        np.savez(path, **self.state_dict())

    @classmethod
    def load(cls, path, venv, **kw):
        d = dict(np.load(path if str(path).endswith(".npz") else str(path) + ".npz"))
        self = cls(venv, norm_obs=bool(d["norm_obs"]), norm_reward=bool(d["norm_reward"]), clip_obs=float(d["clip_obs"]),
                   clip_reward=float(d["clip_reward"]), gamma=float(d["gamma"]), epsilon=float(d["epsilon"]), **kw)
        self.load_state_dict(d)
        return self

    def close(self):
        self.vn.close()
        self.venv.close()
