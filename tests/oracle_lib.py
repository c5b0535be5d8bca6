"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by gym_puzzles_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")


class Layout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_agents", "n_dyn_bodies", "n_fixtures", "n_dyn_fixtures", "max_contacts", "obs_dim", "act_dim",
        "max_episode_steps", "off_goal_contact", "off_bodies", "off_dists", "off_goal", "off_episode_acc",
        "off_aabb", "off_contacts", "state_words")]


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle_capi.cpp", "mrp_env.hpp", "b2core.hpp", "philox.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int]
        for name in ("orc_destroy", "orc_layout", "orc_set_auto_reset", "orc_set_params", "orc_get_params", "orc_reset",
                     "orc_step", "orc_sample_actions", "orc_get_state", "orc_set_state", "orc_stats", "orc_body_mass",
                     "orc_fixture"):
            getattr(L, name).argtypes = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


VARIANTS = {"MultiRobotPuzzle-v0": 0, "MultiRobotPuzzleHeavy-v0": 1, "MultiRobotPuzzle-v2": 2, "MultiRobotPuzzleHeavy-v2": 3,
            "MultiRobotPuzzleSquare-v2": 4}


class OracleBatch:
    def __init__(self, variant, num_envs, seed=17, n_agents=0, env_id_base=0, nthreads=1, max_episode_steps=0):
        if isinstance(variant, str):
            variant = VARIANTS[variant]
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create(variant, n_agents, num_envs, C.c_uint64(seed), C.c_uint64(env_id_base), nthreads))
        if not self.h:
            raise ValueError("orc_create failed")
        self.layout = Layout()
        self.L.orc_layout(self.h, C.byref(self.layout))
        if max_episode_steps > 0:
            self.L.orc_set_max_episode_steps(self.h, max_episode_steps)
        self.N = num_envs
        self.O, self.A, self.SW = self.layout.obs_dim, self.layout.act_dim, self.layout.state_words

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    __del__ = close

    def set_auto_reset(self, on):
        self.L.orc_set_auto_reset(self.h, int(on))

    def set_params(self, p9):
        p = np.ascontiguousarray(p9, dtype=np.float64)
        self.L.orc_set_params(self.h, _p(p))

    def set_curriculum(self, eps=None, decay_pow=None):
        e = None if eps is None else np.ascontiguousarray(eps, dtype=np.float64)
        d = None if decay_pow is None else np.ascontiguousarray(decay_pow, dtype=np.float64)
        self.L.orc_set_curriculum(self.h, None if e is None else _p(e), None if d is None else _p(d))

    def get_params(self):
        p = np.zeros(9)
        self.L.orc_get_params(self.h, _p(p))
        return p

    def reset(self, mask=None):
        obs = np.zeros((self.N, self.O))
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.orc_reset(self.h, None if m is None else _p(m), _p(obs))
        return obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.N, self.A)
        obs = np.zeros((self.N, self.O))
        rew = np.zeros(self.N)
        done = np.zeros(self.N, dtype=np.uint8)
        trunc = np.zeros(self.N, dtype=np.uint8)
        self.L.orc_step(self.h, _p(a), _p(obs), _p(rew), _p(done), _p(trunc))
        return obs, rew, done, trunc

    def sample_actions(self, step_index):
        a = np.zeros((self.N, self.A), dtype=np.float32)
        self.L.orc_sample_actions(self.h, C.c_uint64(step_index), _p(a))
        return a

    def get_state(self, begin=0, count=None):
        count = self.N - begin if count is None else count
        w = np.zeros((count, self.SW), dtype=np.uint32)
        self.L.orc_get_state(self.h, begin, count, _p(w))
        return w

    def set_state(self, words, begin=0):
        w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1, self.SW)
        self.L.orc_set_state(self.h, begin, w.shape[0], _p(w))

    def stats(self):
        s = np.zeros(8)
        self.L.orc_stats(self.h, _p(s))
        return dict(zip(("episodes", "successes", "truncations", "sum_return", "sum_len", "toi_events", "toi_calls", "pos_iters"), s))

    def body_mass(self, body):
        o = np.zeros(6, dtype=np.float32)
        self.L.orc_body_mass(self.h, body, _p(o))
        return o

    def fixture(self, f):
        o = np.zeros(33, dtype=np.float32)
        n = self.L.orc_fixture(self.h, f, _p(o))
        return n, o[:16].reshape(8, 2)[:n], o[16:32].reshape(8, 2)[:n], float(o[32])


class StateView:
    """Decode canonical state words (include/mrp_state.h)."""

    def __init__(self, layout, words):
        self.l, self.w = layout, np.ascontiguousarray(words, dtype=np.uint32).reshape(-1, layout.state_words)

    @property
    def bodies(self):
        l = self.l
        return self.w[:, l.off_bodies:l.off_bodies + 6 * l.n_dyn_bodies].view(np.float32).reshape(-1, l.n_dyn_bodies, 6)

    @property
    def n_contacts(self):
        return self.w[:, 3].astype(np.int32)

    @property
    def goal_contact(self):
        l = self.l
        return self.w[:, l.off_goal_contact:l.off_goal_contact + l.n_agents].astype(np.int32)

    @property
    def dists(self):
        l = self.l
        return np.ascontiguousarray(self.w[:, l.off_dists:l.off_dists + 2 * (l.n_agents + 1)]).view(np.float64)

    @property
    def aabbs(self):
        l = self.l
        return self.w[:, l.off_aabb:l.off_aabb + 4 * l.n_dyn_fixtures].view(np.float32).reshape(-1, l.n_dyn_fixtures, 4)

    @property
    def contacts(self):
        l = self.l
        return self.w[:, l.off_contacts:l.off_contacts + 14 * l.max_contacts].reshape(-1, l.max_contacts, 14)

    def contact_table(self, e):
        """list of dicts for env e in world-list order"""
        out = []
        c = self.contacts[e]
        for k in range(int(self.n_contacts[e])):
            w0 = int(c[k, 0])
            f = c[k].view(np.float32)
            out.append(dict(fA=w0 & 0xff, fB=(w0 >> 8) & 0xff, touching=(w0 >> 16) & 1, type=(w0 >> 17) & 1,
                            pointCount=(w0 >> 18) & 3, keys=int(c[k, 1]), localNormal=f[2:4].copy(), localPoint=f[4:6].copy(),
                            p0=f[6:10].copy(), p1=f[10:14].copy()))
        return out
