"""VectorEnv on the device: zero-copy tensors, auto-reset semantics, statistics, large-batch properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_vector_env_zero_copy_and_auto_reset():
    import torch
    import gym_puzzles_b200 as gp

    env = gp.VectorEnv("MultiRobotPuzzle-v0", 4096, device="cuda:0", seed=3, max_episode_steps=7)
    obs = env.reset()
    assert obs.is_cuda and obs.shape == (4096, 28) and obs.dtype == torch.float32
    assert obs.data_ptr() == env.handle.buffers.obs_dev          # aliases the library buffer
    first = obs.clone()
    dones = 0
    for t in range(21):
        a = torch.rand((4096, 6), device="cuda:0") * 2 - 1
        obs, rew, done, info = env.step(a)
        dones += int(done.sum())
        assert rew.shape == (4096,) and done.dtype == torch.bool
        if t == 6:
            # TimeLimit hit for everyone not already done: obs row is the next episode's first observation
            # (envs that completed earlier have a shifted counter)
            assert float(done.float().mean()) > 0.95 and int(info["TimeLimit.truncated"].sum()) >= 3800
            assert not torch.equal(obs, first)
    st = env.episode_stats(reset=True)
    assert st["episodes"] == dones and st["episodes"] >= 3 * 4096
    assert env.episode_stats()["episodes"] == 0
    # owned action buffer path + on-device synthetic actions
    env.sample_actions(step_index=5)
    assert float(env.actions.abs().max()) <= 1.0 and float(env.actions.std()) > 0.5
    env.step()
    env.close()


def test_gym_style_env_on_gpu_matches_oracle_obs():
    import gym_puzzles_b200 as gp
    from oracle_lib import OracleBatch

    env = gp.make("MultiRobotPuzzleHeavy-v0")
    env.seed(17)
    o = OracleBatch("MultiRobotPuzzleHeavy-v0", 1, seed=17)
    assert np.allclose(env.reset(), o.reset()[0].astype(np.float32), rtol=1e-5, atol=1e-4)
    for t in range(20):
        a = o.sample_actions(t)
        obs, r, d, info = env.step(a[0])
        oo, ro, do, _ = o.step(a)
        assert np.allclose(obs, oo[0].astype(np.float32), rtol=1e-5, atol=1e-4) and d == bool(do[0])
        assert r == pytest.approx(ro[0], rel=1e-4, abs=1e-3)


def test_full_size_properties_heavy_v0():
    """BASELINE.json configs[2] at full size (1,048,576 envs on one GPU): size-independent properties —
    determinism (two handles, same seed => identical tensors), finite state, bodies inside the arena,
    contact flags only where robots are near the block."""
    import torch
    import gym_puzzles_b200 as gp

    N = 1048576
    a = gp.VectorEnv("MultiRobotPuzzleHeavy-v0", N, seed=17)
    b = gp.VectorEnv("MultiRobotPuzzleHeavy-v0", N, seed=17)
    assert torch.equal(a.reset(), b.reset())
    for t in range(12):
        a.sample_actions(t); b.sample_actions(t)
        oa, ra, da, _ = a.step()
        ob, rb, db, _ = b.step()
    assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
    assert bool(torch.isfinite(oa).all()) and bool(torch.isfinite(ra).all())
    # 8 block vertices in px stay inside the 640x480 viewport walls (+ skin)
    verts = oa[:, 24:40]
    assert float(verts.min()) > 29.0 and float(verts[:, 0::2].max()) < 611.0 and float(verts[:, 1::2].max()) < 451.0
    # contact flag implies the robot is within reach of the block: centre distance <= block half-diagonal + robot radius (px)
    dist_px, flag = oa[:, 2:20:4], oa[:, 3:20:4]
    assert float((dist_px * flag).max()) < 4.3 * 30   # 3.354 (farthest block vertex) + 0.79 (robot radius) + skin + one-step lag
    assert 0.0 < float(flag.mean()) < 0.5
    assert a.episode_stats()["overflow"] == 0
    a.close(); b.close()
