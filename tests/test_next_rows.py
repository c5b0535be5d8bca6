"""SURVEY.md §8f "next" rows: SB3-style VecEnv adapter (terminal observation, Monitor info), per-env curriculum
vectors, host render bridge.  CPU tests run through the host build of the kernel source; `-m gpu` repeats the
device-specific parts on the product library."""
import ctypes as C

import numpy as np
import pytest

import gym_puzzles_b200 as gp
from gym_puzzles_b200 import abi
from oracle_lib import OracleBatch


def _emu():
    from emu_lib import emu_lib
    return emu_lib()


def _check_sb3_adapter(lib, env_id="MultiRobotPuzzleHeavy-v0", n=48, cap=15, steps=50):
    """against the oracle stepped WITHOUT auto-reset semantics visible: the oracle's own obs before reset is what the
    adapter must report as terminal_observation"""
    kw = {}
    if lib is not None:
        from emu_lib import host_buffers
        kw = {"_lib": lib, "_buffers": host_buffers}
    venv = gp.SB3VecEnv(env_id, n, seed=7, max_episode_steps=cap, **kw)
    o = OracleBatch(env_id, n, seed=7, max_episode_steps=cap)
    o_term = OracleBatch(env_id, n, seed=7, max_episode_steps=cap)   # same rollout with auto-reset off at the done step
    obs = venv.reset()
    assert obs.shape == (n, venv.observation_space.shape[0]) and obs.dtype == np.float32
    assert np.array_equal(obs, o.reset().astype(np.float32))
    o_term.reset()
    ep_ret, ep_len, seen = np.zeros(n), np.zeros(n, int), 0
    for t in range(steps):
        a = o.sample_actions(t)
        before = o.get_state()
        oobs, orew, odone, otr = o.step(a)
        obs, rew, dones, infos = venv.step(a)
        assert np.array_equal(obs, oobs.astype(np.float32)) and np.array_equal(dones, odone.astype(bool))
        assert len(infos) == n
        ep_ret += orew.astype(np.float32)
        ep_len += 1
        if dones.any():
            # terminal observation: replay the step from the pre-step state with auto-reset disabled
            o_term.set_state(before)
            o_term.set_auto_reset(False)
            tobs, _, tdone, _ = o_term.step(a)
            assert np.array_equal(tdone, odone)
        for i in range(n):
            if dones[i]:
                seen += 1
                info = infos[i]
                assert np.array_equal(info["terminal_observation"], tobs[i].astype(np.float32))
                assert info["TimeLimit.truncated"] == bool(otr[i])
                assert info["episode"]["l"] == ep_len[i] and info["episode"]["t"] >= 0
                assert abs(info["episode"]["r"] - ep_ret[i]) <= 1e-3 * max(1.0, abs(ep_ret[i]))
                ep_ret[i], ep_len[i] = 0.0, 0
            else:
                assert infos[i] == {}
    assert seen >= 2 * n
    venv.env_method("set_reward_params", agentDelta=3.0)
    assert venv.get_attr("weight_deltaAgent")[0] == 3.0
    venv.close()


def test_sb3_adapter_kernel_source():
    _check_sb3_adapter(_emu())


@pytest.mark.gpu
def test_sb3_adapter_gpu():
    _check_sb3_adapter(None, n=512)


def _check_curriculum(lib, n=64):
    """per-env scaled_epsilon / decay_pow: env i uses ITS tolerance and decay (mrp02:227-233, 565-582)"""
    env_id = "MultiRobotPuzzle-v2"
    h = abi.Handle(env_id, n, seed=5, max_episode_steps=30, lib=lib)
    o = OracleBatch(env_id, n, seed=5, max_episode_steps=30)
    e_ptr, d_ptr = h.enable_curriculum()
    eps = 0.1 * (2 - np.arange(n) / n)               # update_goal(epoch=i, nb_epochs=n) per env
    eps[::4] = 5.0                                   # a tolerance so wide that these envs complete at once
    dec = 0.97 ** (-np.arange(n, dtype=np.float64))  # update_params(timestep=i, decay=0.97) per env
    if lib is not None and lib.backend.startswith("host"):
        C.memmove(e_ptr, eps.ctypes.data, eps.nbytes)
        C.memmove(d_ptr, dec.ctypes.data, dec.nbytes)
    else:
        import torch
        from gym_puzzles_b200.vector_env import _wrap
        _wrap(torch, e_ptr, (n,), "<f8", h, torch.device("cuda:0")).copy_(torch.from_numpy(eps))
        _wrap(torch, d_ptr, (n,), "<f8", h, torch.device("cuda:0")).copy_(torch.from_numpy(dec))
        torch.cuda.synchronize()
    o.set_curriculum(eps, dec)
    assert np.array_equal(h.reset_host(), o.reset().astype(np.float32))
    done_seen = 0
    for t in range(12):
        a = o.sample_actions(t)
        oobs, orew, odone, otr = o.step(a)
        obs, rew, done, trunc = h.step_host(a)
        assert np.array_equal(obs, oobs.astype(np.float32)) and np.array_equal(done, odone) and np.array_equal(trunc, otr)
        assert np.allclose(rew, orew, rtol=1e-6, atol=1e-6)
        assert np.array_equal(obs[:, -1], eps.astype(np.float32))       # the obs carries each env's own epsilon (mrp02:531-532)
        done_seen += int(done[::4].sum())
    assert done_seen >= n // 4
    h.close()


def test_curriculum_vectors_kernel_source():
    _check_curriculum(_emu())


@pytest.mark.gpu
def test_curriculum_vectors_gpu():
    _check_curriculum(None, n=2048)


@pytest.mark.gpu
def test_vector_env_curriculum_api():
    import torch
    env = gp.VectorEnv("MultiRobotPuzzleHeavy-v2", 256, seed=1)
    env.update_goal(torch.arange(256), 256)
    env.update_params(torch.arange(256), 0.99)
    assert torch.allclose(env.scaled_epsilon.cpu(), 0.1 * (2 - torch.arange(256, dtype=torch.float64) / 256))
    assert torch.allclose(env.decay_pow.cpu(), 0.99 ** (-torch.arange(256, dtype=torch.float64)))
    env.reset()
    obs, *_ = env.step(torch.zeros(256, 4, device="cuda"))
    assert torch.equal(obs[:, -1].cpu(), env.scaled_epsilon.float().cpu())
    env.close()


def test_render_bridge_scene_and_image():
    from gym_puzzles_b200 import render
    for env_id in ("MultiRobotPuzzle-v0", "MultiRobotPuzzleHeavy-v0", "MultiRobotPuzzle-v2"):
        h = abi.Handle(env_id, 2, seed=3, lib=_emu())
        obs = h.reset_host()
        sc = render.scene(h, 1)
        kinds = [k for k, _ in sc["polygons"]]
        n = h.layout.n_agents
        per_agent = 3 if env_id.endswith("v2") else 1
        assert kinds.count("wall") == 4 and kinds.count("block") == 2 and kinds.count("agent") == n * per_agent
        # the 8 block vertices of the observation (bar then stem, world coords * SCALE / ratio) are the polygon corners
        k = sc["scale"] if env_id.endswith("v0") else sc["scale"] / sc["viewport"][0]
        verts = obs[1, -16:] if env_id.endswith("v0") else obs[1, -17:-1]
        bar = [p for kd, p in sc["polygons"] if kd == "block"][1]
        assert np.allclose(np.asarray(bar).ravel() * k, verts[:8], rtol=1e-4, atol=1e-3 * k)
        img = render.rgb_array(h, 1, downsample=2)
        vw, vh = sc["viewport"]
        assert img.shape == (vh // 2, vw // 2, 3) and img.dtype == np.uint8
        assert (img == 127).all(axis=2).sum() > 20 and (img == 255).all(axis=2).sum() > 10     # block grey, agents white
        h.close()
    env = gp.make("MultiRobotPuzzle-v0", _lib=_emu())
    env.reset()
    assert env.render(mode="rgb_array").shape == (480, 640, 3)


def _obs_v3_numpy(env_id, words, layout):
    """restatement of the reference's experimental -v3 observation (gym_puzzles/envs/core.py:289-350) from canonical state"""
    from oracle_lib import StateView
    sv = StateView(layout, words)
    n = layout.n_agents
    W, H = 640 / 30.0, 480 / 30.0
    ws, hs = W / 2, H / 2
    s = 1.0 if "Heavy" in env_id else 2.0
    verts = [(-3 / s, 0.0), (3 / s, 0.0), (3 / s, 2 / s), (-3 / s, 2 / s), (-1 / s, -2 / s), (1 / s, -2 / s), (1 / s, 0.0), (-1 / s, 0.0)]  # bar, stem
    lc = np.float32(0.5 if "Heavy" in env_id else 0.25)
    out = np.zeros((len(sv.w), 4 * n + 19))
    for e in range(len(sv.w)):
        bod = sv.bodies[e].astype(np.float64)
        bx, by, brot = (bod[0, 0] - ws) / ws, (bod[0, 1] - hs) / ws, bod[0, 2] % (2 * np.pi)
        o = []
        for i in range(n):
            ax, ay = (bod[1 + i, 0] - ws) / ws, (bod[1 + i, 1] - hs) / ws
            o += [bx - ax, by - ay, bod[1 + i, 2] % (2 * np.pi), float(sv.goal_contact[e, i])]
        gx, gy = np.ascontiguousarray(sv.w[e, layout.off_goal:layout.off_goal + 4]).view(np.float64)
        o += [(gx - 320) / 320 - bx, (gy - 240) / 320 - by, 0.0 - brot]
        a = np.float32(sv.bodies[e, 0, 2])
        c, sn = np.float32(np.cos(np.float64(a))), np.float32(np.sin(np.float64(a)))
        px = sv.bodies[e, 0, 0] - (c * np.float32(0) - sn * lc)          # body origin = c - R * localCenter (float32, as Box2D)
        py = sv.bodies[e, 0, 1] - (sn * np.float32(0) + c * lc)
        for vx, vy in verts:
            vx, vy = np.float32(vx), np.float32(vy)
            x = (c * vx - sn * vy) + px
            y = (sn * vx + c * vy) + py
            o += [(np.float64(x) - ws) / ws, (np.float64(y) - hs) / ws]
        out[e] = o
    return out


def _check_obs_v3(lib, device):
    for env_id in ("MultiRobotPuzzle-v0", "MultiRobotPuzzleHeavy-v0"):
        kw = {} if lib is None else {"lib": lib}
        h = abi.Handle(env_id, 64, seed=5, max_episode_steps=30, **kw)
        h.reset_host()
        rng = np.random.default_rng(1)
        for t in range(40):
            h.step_host(rng.uniform(-1, 1, (64, h.act_dim)).astype(np.float32))
        want = _obs_v3_numpy(env_id, h.get_state(), h.layout)
        O3 = 4 * h.layout.n_agents + 19
        if device:
            import torch
            out = torch.empty((64, O3), dtype=torch.float32, device="cuda")
            h.obs_v3(out.data_ptr())
            torch.cuda.synchronize()
            got = out.cpu().numpy()
        else:
            got = np.zeros((64, O3), dtype=np.float32)
            h.obs_v3(got.ctypes.data_as(C.c_void_p))
        assert np.allclose(got, want, rtol=1e-5, atol=2e-6), np.abs(got - want).max()
        assert (got[:, 3::4][:, :h.layout.n_agents] >= 0).all()
        h.close()
    h = abi.Handle("MultiRobotPuzzle-v2", 4, **({} if lib is None else {"lib": lib}))
    with pytest.raises(abi.MrpError):
        h.obs_v3(np.zeros(4 * 27, dtype=np.float32).ctypes.data_as(C.c_void_p))
    h.close()


def test_obs_v3_head_kernel_source():
    _check_obs_v3(_emu(), device=False)


@pytest.mark.gpu
def test_obs_v3_head_gpu():
    _check_obs_v3(None, device=True)
    env = gp.VectorEnv("MultiRobotPuzzleHeavy-v0", 128, seed=2)
    env.reset()
    o3 = env.obs_v3()
    assert tuple(o3.shape) == (128, 39) and bool(o3.isfinite().all())
    env.close()


@pytest.mark.gpu
def test_render_bridge_gpu():
    """render.scene / rgb_array on a CUDA handle: the polygons drawn are the ones the observation's block vertices describe"""
    from gym_puzzles_b200 import render
    for env_id in ("MultiRobotPuzzleHeavy-v0", "MultiRobotPuzzle-v2"):
        h = abi.Handle(env_id, 8, seed=3)
        obs = h.reset_host()
        rng = np.random.default_rng(0)
        for t in range(5):
            obs, *_ = h.step_host(rng.uniform(-1, 1, (8, h.act_dim)).astype(np.float32))
        sc = render.scene(h, 5)
        k = sc["scale"] if env_id.endswith("v0") else sc["scale"] / sc["viewport"][0]
        verts = obs[5, -16:] if env_id.endswith("v0") else obs[5, -17:-1]
        bar = [p for kd, p in sc["polygons"] if kd == "block"][1]
        assert np.allclose(np.asarray(bar).ravel() * k, verts[:8], rtol=1e-4, atol=1e-3 * k)
        img = render.rgb_array(h, 5, downsample=2)
        assert img.shape == (sc["viewport"][1] // 2, sc["viewport"][0] // 2, 3)
        h.close()
