"""Loader of tests/emu/libmrp_emu.so — the kernel source of gym_puzzles_b200/csrc compiled for the host
(-DMRP_HOST_EMU).  Debug/test tool for the GPU-less build container only; the package never loads it."""
import os
import subprocess

from gym_puzzles_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(_HERE, "emu")
_lib = None


def emu_lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", EMU_DIR, "-s"])
        _lib = abi.MrpLib(os.path.join(EMU_DIR, "libmrp_emu.so"))
        assert _lib.backend.startswith("host-emu")
    return _lib


def host_buffers(h, term, n_envs, device):
    """`_buffers` hook of gym_puzzles_b200.SB3VecEnv for the host build: its "device" pointers are host pointers."""
    import ctypes as C

    import numpy as np

    def view(p, shape, ct, dt):
        return np.frombuffer((ct * int(np.prod(shape))).from_address(p), dtype=dt).reshape(shape)

    t_obs = view(term.terminal_obs_dev, (n_envs, h.obs_dim), C.c_float, np.float32)
    t_ret = view(term.episode_return_dev, (n_envs,), C.c_float, np.float32)
    t_len = view(term.episode_length_dev, (n_envs,), C.c_int32, np.int32)
    gather = lambda idx: (t_obs[idx].copy(), t_ret[idx].copy(), t_len[idx].copy())  # noqa: E731
    return (t_obs, t_ret, t_len, gather, np.empty((n_envs, h.obs_dim), np.float32), np.empty(n_envs, np.float32),
            np.empty(n_envs, np.uint8), np.empty(n_envs, np.uint8), np.empty((n_envs, h.act_dim), np.float32))


def emu_device_stepper(handle):
    """step(actions) through mrp_step of the host build (its "device" buffers are host memory): exercises the code path of
    the device-resident entry point (task-free / task-owning k_post groups and their TOI queues) on the CPU."""
    import ctypes as C

    import numpy as np

    b, N = handle.buffers, handle.num_envs

    def view(p, shape, ct, dt):
        return np.frombuffer((ct * int(np.prod(shape))).from_address(p), dtype=dt).reshape(shape)

    act = view(b.action_dev, (N, handle.act_dim), C.c_float, np.float32)
    obs = view(b.obs_dev, (N, handle.obs_dim), C.c_float, np.float32)
    rew = view(b.reward_dev, (N,), C.c_float, np.float32)
    done = view(b.done_dev, (N,), C.c_uint8, np.uint8)
    trunc = view(b.trunc_dev, (N,), C.c_uint8, np.uint8)

    def step(a):
        act[...] = a
        handle.step()
        return obs.copy(), rew.copy(), done.copy(), trunc.copy()

    return step
