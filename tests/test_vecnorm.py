"""VecNormalize kernels (csrc/mrp_vecnorm.cu) against a numpy restatement of Stable-Baselines3's VecNormalize /
RunningMeanStd (the wrapper the reference's trainer uses, train/train.py:82).  The CPU tests run the host build of the
kernel source; the `-m gpu` test runs the sm_100a kernels on device tensors."""
import numpy as np
import pytest

from gym_puzzles_b200.vec_normalize import VecNormHandle, VN_EXPORTS


class RunningMeanStd:
    """stable_baselines3.common.running_mean_std.RunningMeanStd (published algorithm, float64)."""

    def __init__(self, shape=(), epsilon=1e-4):
        self.mean, self.var, self.count = np.zeros(shape, np.float64), np.ones(shape, np.float64), epsilon

    def update(self, arr):
        arr = np.asarray(arr, dtype=np.float64)
        bm, bv, bc = arr.mean(axis=0), arr.var(axis=0), arr.shape[0]
        delta = bm - self.mean
        tot = self.count + bc
        new_mean = self.mean + delta * bc / tot
        m2 = self.var * self.count + bv * bc + np.square(delta) * self.count * bc / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


class NumpyVecNormalize:
    """stable_baselines3.common.vec_env.VecNormalize.step_wait / reset (published algorithm)."""

    def __init__(self, n, o, clip_obs=10.0, clip_reward=10.0, gamma=0.99, epsilon=1e-8):
        self.obs_rms, self.ret_rms = RunningMeanStd((o,)), RunningMeanStd(())
        self.returns = np.zeros(n)
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon, self.training = clip_obs, clip_reward, gamma, epsilon, True

    def norm_obs(self, obs):
        return np.clip((obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon), -self.clip_obs, self.clip_obs).astype(np.float32)

    def reset(self, obs):
        self.returns[:] = 0
        if self.training:
            self.obs_rms.update(obs)
        return self.norm_obs(obs)

    def step(self, obs, rew, done):
        if self.training:
            self.obs_rms.update(obs)
        o = self.norm_obs(obs)
        if self.training:
            self.returns = self.returns * self.gamma + rew
            self.ret_rms.update(self.returns)
        r = np.clip(rew / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward).astype(np.float32)
        self.returns[done.astype(bool)] = 0
        return o, r


def _run(make_handle, to_dev, to_host, ptr, N=3000, O=40, steps=12):
    rng = np.random.default_rng(0)
    ref = NumpyVecNormalize(N, O)
    vn = make_handle(N, O)
    scale = rng.uniform(0.1, 300.0, O).astype(np.float32)
    shift = rng.uniform(-200, 500, O).astype(np.float32)

    def batch():
        return (rng.standard_normal((N, O)).astype(np.float32) * scale + shift), (rng.standard_normal(N) * 30 - 20).astype(np.float32), \
            (rng.uniform(size=N) < 0.1).astype(np.uint8)

    obs, _, _ = batch()
    want = ref.reset(obs)
    d_obs, d_out = to_dev(obs), to_dev(np.zeros_like(obs))
    vn.reset_returns()
    vn.moments(ptr(d_obs))
    vn.apply(ptr(d_obs), None, None, ptr(d_out))
    np.testing.assert_allclose(to_host(d_out), want, rtol=2e-5, atol=2e-5)
    for t in range(steps):
        if t == steps - 3:          # evaluation mode: statistics frozen (train/test.py:67)
            ref.training = False
            vn.set_training(False)
        obs, rew, done = batch()
        term = rng.standard_normal((N, O)).astype(np.float32) * scale + shift
        want_o, want_r = ref.step(obs, rew, done)
        want_t = np.where(done[:, None].astype(bool), ref.norm_obs(term), term)
        d_obs, d_rew, d_done, d_term = to_dev(obs), to_dev(rew), to_dev(done), to_dev(term)
        d_out, d_rout = to_dev(np.zeros_like(obs)), to_dev(np.zeros_like(rew))
        vn.moments(ptr(d_obs), ptr(d_rew))
        vn.apply(ptr(d_obs), ptr(d_rew), ptr(d_done), ptr(d_out), ptr(d_rout), ptr(d_term))
        np.testing.assert_allclose(to_host(d_out), want_o, rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(to_host(d_rout), want_r, rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(to_host(d_term), want_t, rtol=2e-5, atol=2e-5)
    s = vn.get_stats()
    np.testing.assert_allclose(s[:O], ref.obs_rms.mean, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(s[O + 1:2 * O + 1], ref.obs_rms.var, rtol=1e-8)
    np.testing.assert_allclose([s[O], s[2 * O + 1]], [ref.ret_rms.mean, ref.ret_rms.var], rtol=1e-8)
    np.testing.assert_allclose(s[2 * O + 2:], [ref.obs_rms.count, ref.ret_rms.count], rtol=1e-12)
    # save / load round trip
    vn2 = make_handle(N, O)
    vn2.set_stats(s)
    assert np.array_equal(vn2.get_stats(), s)
    return vn


def test_vecnorm_kernel_source_matches_sb3_algorithm():
    from emu_lib import emu_lib
    for O in (40, 39, 66):          # float4 path, ragged rows, three columns per lane
        _run(lambda n, o: VecNormHandle(n, o, lib=emu_lib()), lambda a: a.copy(), lambda a: a, lambda a: a.ctypes.data, O=O)


def test_vecnorm_symbols_exported():
    from gym_puzzles_b200 import abi
    import ctypes, os
    if not os.path.exists(abi.LIB_PATH):
        pytest.skip("product library not built")
    L = ctypes.CDLL(abi.LIB_PATH)
    assert all(hasattr(L, s) for s in VN_EXPORTS)


@pytest.mark.gpu
def test_vecnorm_gpu_matches_sb3_algorithm():
    import torch
    for O, N in ((40, 70001), (39, 5000), (66, 3000)):
        vn = _run(lambda n, o: VecNormHandle(n, o, device=0), lambda a: torch.from_numpy(a).cuda(), lambda t: t.cpu().numpy(),
                  lambda t: t.data_ptr(), N=N, O=O)
        assert vn.launch_count > 0


@pytest.mark.gpu
def test_vecnormalize_wrapper_over_vector_env():
    import torch
    import gym_puzzles_b200 as gp
    from gym_puzzles_b200.vec_normalize import VecNormalize
    env = VecNormalize(gp.VectorEnv("MultiRobotPuzzleHeavy-v0", 4096, seed=3, max_episode_steps=20))
    ref = NumpyVecNormalize(4096, 40)
    obs = env.reset()
    np.testing.assert_allclose(obs.cpu().numpy(), ref.reset(env.get_original_obs().cpu().numpy()), rtol=2e-5, atol=2e-5)
    for t in range(30):
        env.venv.sample_actions(t)
        obs, rew, done, info = env.step()
        wo, wr = ref.step(env.get_original_obs().cpu().numpy(), env.get_original_reward().cpu().numpy(), done.cpu().numpy())
        np.testing.assert_allclose(obs.cpu().numpy(), wo, rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(rew.cpu().numpy(), wr, rtol=2e-5, atol=2e-5)
    sd = env.state_dict()
    np.testing.assert_allclose(sd["obs_rms.mean"], ref.obs_rms.mean, rtol=1e-8, atol=1e-8)
    env.close()


def test_vecnorm_rank_sum_equals_global_batch():
    """Multi-GPU VecNormalize: every rank computes shifted moments of its shard, the accumulator vectors are summed
    (one NCCL all-reduce on the device; here the sum is done by hand on the host build) and every rank then holds the
    statistics of ONE VecNormalize over the global batch."""
    import ctypes as C
    from emu_lib import emu_lib
    N, O, R = 1200, 40, 3
    rng = np.random.default_rng(3)
    ranks = [VecNormHandle(N // R, O, lib=emu_lib()) for _ in range(R)]
    whole = VecNormHandle(N, O, lib=emu_lib())

    def accum(vn):
        p, n = vn.accum()
        return np.frombuffer((C.c_double * n).from_address(p), dtype=np.float64)

    for t in range(6):
        obs = (rng.standard_normal((N, O)) * 40 + 7).astype(np.float32)
        rew = (rng.standard_normal(N) * 3).astype(np.float32)
        done = (rng.uniform(size=N) < 0.2).astype(np.uint8)
        out_w, rout_w = np.zeros_like(obs), np.zeros_like(rew)
        whole.moments(obs.ctypes.data, rew.ctypes.data)
        whole.apply(obs.ctypes.data, rew.ctypes.data, done.ctypes.data, out_w.ctypes.data, rout_w.ctypes.data)
        shards = np.split(np.arange(N), R)
        for vn, ix in zip(ranks, shards):
            o, r = np.ascontiguousarray(obs[ix]), np.ascontiguousarray(rew[ix])
            vn.moments(o.ctypes.data, r.ctypes.data)
        total = sum(accum(vn).copy() for vn in ranks)          # the all-reduce
        for vn in ranks:
            accum(vn)[:] = total
        for vn, ix in zip(ranks, shards):
            o, r, d = (np.ascontiguousarray(a[ix]) for a in (obs, rew, done))
            oo, ro = np.zeros_like(o), np.zeros_like(r)
            vn.apply(o.ctypes.data, r.ctypes.data, d.ctypes.data, oo.ctypes.data, ro.ctypes.data)
            np.testing.assert_allclose(oo, out_w[ix], rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(ro, rout_w[ix], rtol=1e-5, atol=1e-5)
    for vn in ranks:
        np.testing.assert_allclose(vn.get_stats(), whole.get_stats(), rtol=1e-10, atol=1e-12)
