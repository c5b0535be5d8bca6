"""ctypes binding of the C-ABI in include/mrp_b200.h (libmrp_b200.so).

The product library is CUDA-only: ``load()`` raises if the in-tree shared object is missing
and ``Handle`` creation raises if no CUDA device is present — there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmrp_b200.so")

VARIANTS = {
    "MultiRobotPuzzle-v0": 0,
    "MultiRobotPuzzleHeavy-v0": 1,
    "MultiRobotPuzzle-v2": 2,
    "MultiRobotPuzzleHeavy-v2": 3,
    # extension (BASELINE.json configs[4]; not a reference env): three blocks T, L, I that form a square, Heavy-v2 dynamics
    "MultiRobotPuzzleSquare-v2": 4,
}

N_STATS = 16
STAT_NAMES = ("episodes", "done_by_env", "truncated", "sum_return", "sum_return_sq", "sum_length", "env_steps", "overflow",
              "nan_resets", "pairs", "m1", "m2", "vel_flops", "pos_points", "toi_calls", "reserved")

EXPORTS = (
    "mrp_last_error", "mrp_backend", "mrp_create", "mrp_destroy", "mrp_get_layout", "mrp_get_buffers", "mrp_reset",
    "mrp_step", "mrp_step_host", "mrp_reset_host", "mrp_sample_actions", "mrp_get_state", "mrp_set_state",
    "mrp_set_params", "mrp_get_params", "mrp_get_stats", "mrp_set_timing", "mrp_get_timing", "mrp_get_phase_timing", "mrp_launch_count",
    "mrp_enable_terminal_info", "mrp_enable_curriculum", "mrp_obs_v3",
)


class Config(C.Structure):
    _fields_ = [("variant", C.c_int32), ("n_agents", C.c_int32), ("num_envs", C.c_int32), ("device", C.c_int32),
                ("seed", C.c_uint64), ("env_id_base", C.c_uint64), ("auto_reset", C.c_int32),
                ("max_episode_steps", C.c_int32)]


class Layout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_agents", "n_dyn_bodies", "n_fixtures", "n_dyn_fixtures", "max_contacts", "obs_dim", "act_dim",
        "max_episode_steps", "off_goal_contact", "off_bodies", "off_dists", "off_goal", "off_episode_acc",
        "off_aabb", "off_contacts", "state_words")]


class Buffers(C.Structure):
    _fields_ = [("action_dev", C.c_void_p), ("obs_dev", C.c_void_p), ("reward_dev", C.c_void_p),
                ("done_dev", C.c_void_p), ("trunc_dev", C.c_void_p), ("stats_dev", C.c_void_p),
                ("num_envs", C.c_int32), ("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("reserved", C.c_int32)]


class TerminalBuffers(C.Structure):
    _fields_ = [("terminal_obs_dev", C.c_void_p), ("episode_return_dev", C.c_void_p), ("episode_length_dev", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "agentDelta", "agentDistance", "blockDelta", "blockDistance", "puzzleComp", "outOfBounds", "blkOutOfBounds",
        "scaled_epsilon", "decay_pow")]


class MrpError(RuntimeError):
    pass


class MrpLib:
    """A loaded mrp C-ABI library."""

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise MrpError(
                f"{path} not found: build the sm_100a extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                "gym_puzzles_b200 has no CPU fallback.")
        self.path = path
        L = self.lib = C.CDLL(path)
        L.mrp_last_error.restype = C.c_char_p
        L.mrp_backend.restype = C.c_char_p
        L.mrp_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.mrp_destroy.argtypes = [C.c_void_p]
        L.mrp_get_layout.argtypes = [C.c_void_p, C.POINTER(Layout)]
        L.mrp_get_buffers.argtypes = [C.c_void_p, C.POINTER(Buffers)]
        L.mrp_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrp_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrp_step_host.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.mrp_reset_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrp_sample_actions.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.mrp_enable_terminal_info.argtypes = [C.c_void_p, C.POINTER(TerminalBuffers)]
        L.mrp_enable_curriculum.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.mrp_get_state.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.mrp_set_state.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.mrp_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.mrp_get_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.mrp_get_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.mrp_set_timing.argtypes = [C.c_void_p, C.c_int32]
        L.mrp_get_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]
        L.mrp_get_phase_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.mrp_launch_count.argtypes = [C.c_void_p]
        L.mrp_obs_v3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrp_launch_count.restype = C.c_int64

    @property
    def backend(self):
        return self.lib.mrp_backend().decode()

    def check(self, rc, what):
        if rc != 0:
            raise MrpError(f"{what} failed ({rc}): {self.lib.mrp_last_error().decode()}")


_default = None


def load():
    """The product library (CUDA, sm_100a).  Fails loudly when it has not been built."""
    global _default
    if _default is None:
        # MRP_LIB_PATH: another build of the same library (A/B measurements of two builds on one box); still CUDA-only
        _default = MrpLib(os.environ.get("MRP_LIB_PATH", LIB_PATH))
        if not _default.backend.startswith("cuda"):
            raise MrpError(f"{_default.path} is not the CUDA library ({_default.backend}): gym_puzzles_b200 has no CPU fallback")
    return _default


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Handle:
    """Owns one mrp_handle: a shard of `num_envs` environments resident on one GPU."""

    def __init__(self, variant, num_envs, device=0, seed=17, n_agents=0, env_id_base=0, auto_reset=True,
                 max_episode_steps=0, lib=None):
        self.lib = lib if lib is not None else load()
        if isinstance(variant, str):
            if variant not in VARIANTS:
                raise KeyError(f"unknown env id {variant!r}; known: {sorted(VARIANTS)}")
            variant = VARIANTS[variant]
        cfg = Config(variant, n_agents, num_envs, device, seed, env_id_base, 1 if auto_reset else 0, max_episode_steps)
        self.h = C.c_void_p()
        self.lib.check(self.lib.lib.mrp_create(C.byref(cfg), C.byref(self.h)), "mrp_create")
        self.layout = Layout()
        self.lib.check(self.lib.lib.mrp_get_layout(self.h, C.byref(self.layout)), "mrp_get_layout")
        self.buffers = Buffers()
        self.lib.check(self.lib.lib.mrp_get_buffers(self.h, C.byref(self.buffers)), "mrp_get_buffers")
        self.num_envs = num_envs
        self.obs_dim, self.act_dim = self.layout.obs_dim, self.layout.act_dim
        self.state_words = self.layout.state_words
        self.state_bytes = 4 * self.state_words   # canonical record; the device-resident [word][env] state is about as large
        self.variant = variant

    def close(self):
        if getattr(self, "h", None):
            self.lib.lib.mrp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stream-ordered device calls
    def reset(self, mask_ptr=None, stream=None):
        self.lib.check(self.lib.lib.mrp_reset(self.h, mask_ptr, stream), "mrp_reset")

    def step(self, actions_ptr=None, stream=None):
        self.lib.check(self.lib.lib.mrp_step(self.h, actions_ptr, stream), "mrp_step")

    def sample_actions(self, step_index, dst_ptr=None, stream=None):
        self.lib.check(self.lib.lib.mrp_sample_actions(self.h, step_index, dst_ptr, stream), "mrp_sample_actions")

    # ---- host-buffer calls (numpy)
    def step_host(self, actions, obs=None, reward=None, done=None, trunc=None):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.num_envs, self.act_dim)
        obs = np.empty((self.num_envs, self.obs_dim), dtype=np.float32) if obs is None else obs
        reward = np.empty(self.num_envs, dtype=np.float32) if reward is None else reward
        done = np.empty(self.num_envs, dtype=np.uint8) if done is None else done
        trunc = np.empty(self.num_envs, dtype=np.uint8) if trunc is None else trunc
        self.lib.check(self.lib.lib.mrp_step_host(self.h, _ptr(a), _ptr(obs), _ptr(reward), _ptr(done), _ptr(trunc)), "mrp_step_host")
        return obs, reward, done, trunc

    def reset_host(self, mask=None, obs=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        obs = np.empty((self.num_envs, self.obs_dim), dtype=np.float32) if obs is None else obs
        self.lib.check(self.lib.lib.mrp_reset_host(self.h, _ptr(m), _ptr(obs)), "mrp_reset_host")
        return obs

    def get_state(self, begin=0, count=None):
        count = self.num_envs - begin if count is None else count
        w = np.zeros((count, self.state_words), dtype=np.uint32)
        self.lib.check(self.lib.lib.mrp_get_state(self.h, begin, count, _ptr(w)), "mrp_get_state")
        return w

    def set_state(self, words, begin=0):
        w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1, self.state_words)
        self.lib.check(self.lib.lib.mrp_set_state(self.h, begin, w.shape[0], _ptr(w)), "mrp_set_state")

    def set_params(self, **kw):
        p = self.get_params_struct()
        for k, v in kw.items():
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, float(v))
        self.lib.check(self.lib.lib.mrp_set_params(self.h, C.byref(p)), "mrp_set_params")

    def get_params_struct(self):
        p = Params()
        self.lib.check(self.lib.lib.mrp_get_params(self.h, C.byref(p)), "mrp_get_params")
        return p

    def get_params(self):
        p = self.get_params_struct()
        return {n: getattr(p, n) for n, _ in Params._fields_}

    def enable_terminal_info(self):
        """-> TerminalBuffers (device pointers): terminal observation / episode return / length of envs done this step"""
        t = TerminalBuffers()
        self.lib.check(self.lib.lib.mrp_enable_terminal_info(self.h, C.byref(t)), "mrp_enable_terminal_info")
        return t

    def enable_curriculum(self):
        """-> (scaled_epsilon_dev, decay_pow_dev): device pointers of the per-env f64[num_envs] curriculum vectors"""
        e, d = C.c_void_p(), C.c_void_p()
        self.lib.check(self.lib.lib.mrp_enable_curriculum(self.h, C.byref(e), C.byref(d)), "mrp_enable_curriculum")
        return e.value, d.value

    def obs_v3(self, out_ptr, stream=None):
        """normalised MultiRobotPuzzle-v3 observation head (core.py:289-350) into f32[num_envs][4 n + 19] at out_ptr"""
        self.lib.check(self.lib.lib.mrp_obs_v3(self.h, out_ptr, stream), "mrp_obs_v3")

    def stats(self, reset_after=False):
        s = np.zeros(N_STATS, dtype=np.float64)
        self.lib.check(self.lib.lib.mrp_get_stats(self.h, _ptr(s), 1 if reset_after else 0), "mrp_get_stats")
        return dict(zip(STAT_NAMES, s.tolist()))

    def set_timing(self, enable):
        self.lib.check(self.lib.lib.mrp_set_timing(self.h, 1 if enable else 0), "mrp_set_timing")

    def get_timing(self, reset_after=True):
        """-> (total milliseconds spent in the step kernel, number of step-kernel launches)"""
        ms, n = C.c_double(), C.c_int64()
        self.lib.check(self.lib.lib.mrp_get_timing(self.h, C.byref(ms), C.byref(n), 1 if reset_after else 0), "mrp_get_timing")
        return ms.value, n.value

    PHASES = ("k_pre", "k_solve_vel", "k_solve_pos", "k_post", "k_post_events")

    def get_phase_timing(self, reset_after=True):
        """-> {kernel name: accumulated milliseconds} for the five phase kernels of step()"""
        ms = np.zeros(5, dtype=np.float64)
        self.lib.check(self.lib.lib.mrp_get_phase_timing(self.h, _ptr(ms), 1 if reset_after else 0), "mrp_get_phase_timing")
        return dict(zip(self.PHASES, ms.tolist()))

    @property
    def launch_count(self):
        return int(self.lib.lib.mrp_launch_count(self.h))
