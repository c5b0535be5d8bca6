"""Size-independent properties of the hot path (SURVEY.md §4 iii), checked on the kernel source (CPU) and, at
BASELINE.json's full batch size, on the sm_100a library:
  * accumulated normal impulses are >= 0 and friction impulses stay inside the cone |jt| <= mu * jn (b2ContactSolver);
  * rewards / observations are finite, done == (env done | TimeLimit), truncation only at the cap;
  * determinism: the same seed gives the same trajectory; a batch equals the concatenation of its shards."""
import numpy as np
import pytest

from gym_puzzles_b200 import abi
from oracle_lib import StateView

FRICTION = {0: (0.999, 0.2), 1: (0.999, 0.2), 2: (0.01, 0.01), 3: (0.01, 0.01)}   # (block, robot) fixtures; walls 0.2


def _contact_properties(h, words, variant):
    sv = StateView(h.layout, words)
    l = h.layout
    per_agent = 3 if variant >= 2 else 1
    fb, fr = FRICTION[variant]
    fric = np.array([fb, fb] + [fr] * (per_agent * l.n_agents) + [0.2] * 4)
    c = sv.contacts
    nc = sv.n_contacts
    checked, worst = 0, -1.0
    for e in range(c.shape[0]):
        for k in range(int(nc[e])):
            w0 = int(c[e, k, 0])
            pc = (w0 >> 18) & 3
            if pc == 0:
                continue
            mu = np.float32(np.sqrt(np.float32(fric[w0 & 0xff] * fric[(w0 >> 8) & 0xff])))
            f = c[e, k].view(np.float32)
            for p in range(pc):
                jn, jt = f[8 + 4 * p], f[9 + 4 * p]
                assert jn >= 0.0, (e, k, jn)
                # b2ContactSolver clamps friction against the normal impulse of the PREVIOUS normal update (friction is
                # solved first in every sweep), so the cone holds up to the last sweep's change of jn
                worst = max(worst, (abs(jt) - mu * jn) / max(mu * jn, 1e-6))
                assert abs(jt) <= mu * jn * 1.25 + 1e-3, (e, k, jt, mu, jn)
                checked += 1
    return checked, worst


@pytest.mark.parametrize("variant", [1, 2])
def test_contact_impulse_properties_kernel_source(variant):
    from emu_lib import emu_lib
    N = 256
    h = abi.Handle(variant, N, seed=31, max_episode_steps=80, lib=emu_lib())
    h.reset_host()
    rng = np.random.default_rng(5)
    checked = 0
    for t in range(60):
        a = rng.uniform(-1, 1, (N, h.act_dim)).astype(np.float32)
        if variant >= 2:
            a[:, 1::2] = 1.0
        obs, rew, done, trunc = h.step_host(a)
        assert np.isfinite(obs).all() and np.isfinite(rew).all()
        assert not (trunc & ~done).any()
        if t % 10 == 9:
            n, worst = _contact_properties(h, h.get_state(), variant)
            checked += n
            print('cone excess', worst)
    assert checked > 50


@pytest.mark.gpu
def test_full_size_properties_heavy_v0():
    """BASELINE.json configs[2] size: 1,048,576 MultiRobotPuzzleHeavy-v0 envs on one GPU."""
    import torch
    N = 1 << 20
    env_id = "MultiRobotPuzzleHeavy-v0"

    def rollout(handles, steps):
        sums = []
        for h in handles:
            h.reset()
        for t in range(steps):
            for h in handles:
                h.sample_actions(t)
                h.step()
        torch.cuda.synchronize()

    from gym_puzzles_b200.vector_env import _wrap
    dev = torch.device("cuda:0")

    def obs_of(h):
        return _wrap(torch, h.buffers.obs_dev, (h.num_envs, h.obs_dim), "<f4", h, dev)

    a = abi.Handle(env_id, N, seed=9, max_episode_steps=25)
    rollout([a], 40)
    oa = obs_of(a).clone()
    assert torch.isfinite(oa).all()
    st = a.stats()
    assert st["episodes"] >= N and st["overflow"] == 0 and st["truncated"] + st["done_by_env"] == st["episodes"]
    # contact-impulse properties on a sample of the batch
    assert _contact_properties(a, a.get_state(0, 4096), 1)[0] > 1000
    a.close()
    # determinism: same seed, same trajectory (bitwise)
    b = abi.Handle(env_id, N, seed=9, max_episode_steps=25)
    rollout([b], 40)
    assert torch.equal(obs_of(b), oa)
    b.close()
    # sharding invariance at scale: two shards of N/2 with global env ids == one batch of N
    s0 = abi.Handle(env_id, N // 2, seed=9, max_episode_steps=25, env_id_base=0)
    s1 = abi.Handle(env_id, N // 2, seed=9, max_episode_steps=25, env_id_base=N // 2)
    rollout([s0, s1], 40)
    assert torch.equal(torch.cat([obs_of(s0), obs_of(s1)]), oa)
    s0.close(); s1.close()
