"""One env.step from IDENTICAL states (BASELINE.json north_star), three ways:

 * the reference's own Python (unmodified mrp00 / mrp02 over tests/refshim) vs the oracle      — build container only;
 * the host build of the kernel source vs the oracle                                           — CPU;
 * the sm_100a library vs the oracle, through mrp_step_host and through the device path        — `-m gpu`.

Covers the v2 termination branches no rollout reaches (robot / block out of bounds, mrp02:279-295,552-563, with
update_params decay != 1) besides ordinary mid-rollout states of all four ids.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim"))
import harness  # noqa: E402
import state_cases  # noqa: E402
from gym_puzzles_b200 import abi  # noqa: E402
from oracle_lib import OracleBatch, StateView  # noqa: E402
from parity_util import compare_states, device_stepper  # noqa: E402

IDS = list(harness.REGISTRY)
V2 = [i for i in IDS if i.endswith("v2")]
needs_reference = pytest.mark.skipif(not harness.available(), reason="reference checkout not present")
DECAY = (7, 0.93)     # update_params(timestep, decay): shaped penalties = weight * decay ** (-timestep)


def _close(a, b, amplify=1.0):
    return np.allclose(a, b, rtol=4e-15, atol=1e-13 * amplify)


def _oracle_one(env_id, state, action, params=None, n_agents=0, gid=0, cap=0):
    o = OracleBatch(env_id, 1, seed=state_cases.SEED, env_id_base=gid, max_episode_steps=cap, n_agents=n_agents)
    o.set_auto_reset(False)
    if params is not None:
        o.set_params(params)
    o.set_state(state[None])
    out = o.step(action[None])
    w = o.get_state()
    o.close()
    return out, w


def _actions(env_id, n, n_agents=0):
    o = OracleBatch(env_id, n, seed=5, n_agents=n_agents)
    a = o.sample_actions(3)
    o.close()
    return a


def _decayed_params(env_id, n_agents=0):
    o = OracleBatch(env_id, 1, n_agents=n_agents)
    p = o.get_params()
    o.close()
    p[8] = DECAY[1] ** (-DECAY[0])
    return p


# ---------------------------------------------------------------------------------------- reference Python vs oracle
def _reference_one(env_id, layout, state, action, decay=None, num_agents=None):
    r = harness.ReferenceEnv(env_id, seed=state_cases.SEED, gid=0, num_agents=num_agents)
    r.reset()
    if decay is not None:
        r.env.update_params(*decay)
    r.set_state(layout, state)
    with r._feeds():
        obs, rew, done, info = r.env.step(np.asarray(action, dtype=np.float32).astype(np.float64))
    return r, np.asarray(obs, dtype=np.float64), float(rew), bool(done)


@needs_reference
@pytest.mark.parametrize("env_id", IDS)
def test_reference_python_one_step_from_rollout_states(env_id):
    states = state_cases.rollout_states(env_id, 10)
    acts = _actions(env_id, len(states))
    touching = 0
    for s, a in zip(states, acts):
        (oobs, orew, odone, otr), w = _oracle_one(env_id, s, a)
        lay = OracleBatch(env_id, 1).layout
        r, obs, rew, done = _reference_one(env_id, lay, s, a)
        assert _close(obs, oobs[0]) and _close(rew, orew[0], 1e3)
        assert done == bool(odone[0])
        sv = StateView(lay, w)
        assert r.goal_contacts == [bool(x) for x in sv.goal_contact[0]]
        assert np.array_equal(r.body_rows().view(np.uint32), sv.bodies[0].view(np.uint32))
        touching += int(((StateView(lay, s[None]).contacts[0][:, 0] >> 16) & 1).sum())
    assert touching > 0     # the states carry touching contacts with warm-start impulses


@needs_reference
@pytest.mark.parametrize("env_id", V2)
@pytest.mark.parametrize("decay", [None, DECAY])
def test_reference_python_v2_out_of_bounds(env_id, decay):
    """mrp02:552-563: robot OOB -> -outOfBounds * decay**(-t), done; else block OOB -> -blkOutOfBounds * decay**(-t), done."""
    states, kinds = state_cases.oob_states(env_id)
    acts = _actions(env_id, len(states))
    params = _decayed_params(env_id) if decay else None
    lay = OracleBatch(env_id, 1).layout
    scale = DECAY[1] ** (-DECAY[0]) if decay else 1.0
    for s, kind, a in zip(states, kinds, acts):
        (oobs, orew, odone, otr), w = _oracle_one(env_id, s, a, params=params)
        r, obs, rew, done = _reference_one(env_id, lay, s, a, decay=decay or (0, 1.0))
        assert done and odone[0] == 1 and otr[0] == 0, kind
        assert _close(obs, oobs[0]) and _close(rew, orew[0], 1e3), kind
        assert np.array_equal(r.body_rows().view(np.uint32), StateView(lay, w).bodies[0].view(np.uint32))
        # the penalty itself: shaping terms are O(1), the penalties 1000 / 100 (x decay**-t)
        want = -(100.0 if kind == "block" else 1000.0) * scale
        assert abs(rew - want) < 15.0, (kind, rew, want)
        assert "Out Of Bounds" in r.env.done_status and ("Agent" in r.env.done_status) == (kind != "block")


@needs_reference
def test_reference_python_v2_out_of_bounds_more_agents():
    env_id = "MultiRobotPuzzle-v2"
    states, kinds = state_cases.oob_states(env_id, n_agents=4)
    acts = _actions(env_id, len(states), n_agents=4)
    lay = OracleBatch(env_id, 1, n_agents=4).layout
    for s, kind, a in list(zip(states, kinds, acts))[::3]:
        (oobs, orew, odone, otr), w = _oracle_one(env_id, s, a, n_agents=4)
        r, obs, rew, done = _reference_one(env_id, lay, s, a, decay=(0, 1.0), num_agents=4)
        assert done and odone[0] == 1
        assert _close(obs, oobs[0]) and _close(rew, orew[0], 1e3)


# ---------------------------------------------------------------------------------------- kernel source / CUDA vs oracle
def _batch_case(env_id, n_agents=0, with_oob=True):
    """states + actions + params of one batched single-step comparison"""
    states = [state_cases.rollout_states(env_id, 24, n_agents=n_agents)]
    if with_oob and env_id.endswith("v2"):
        states.append(state_cases.oob_states(env_id, n_agents=n_agents)[0])
    states = np.concatenate(states)
    return states, _actions(env_id, len(states), n_agents=n_agents)


def _abi_vs_oracle(env_id, lib=None, device_path=False, n_agents=0, decay=False, auto_reset=True):
    states, acts = _batch_case(env_id, n_agents=n_agents)
    N = len(states)
    kw = {} if lib is None else {"lib": lib}
    h = abi.Handle(env_id, N, seed=state_cases.SEED, n_agents=n_agents, auto_reset=auto_reset, **kw)
    o = OracleBatch(env_id, N, seed=state_cases.SEED, nthreads=4, n_agents=n_agents)
    o.set_auto_reset(auto_reset)
    h.reset_host()
    if decay:
        p = _decayed_params(env_id, n_agents=n_agents)
        o.set_params(p)
        h.set_params(decay_pow=float(p[8]))
    o.set_state(states)
    h.set_state(states)
    obs_o, r_o, d_o, t_o = o.step(acts)
    step = device_stepper(h) if device_path else h.step_host
    obs_h, r_h, d_h, t_h = step(acts)
    assert np.array_equal(d_o, d_h) and np.array_equal(t_o, t_h)
    assert np.allclose(obs_o.astype(np.float32), obs_h, rtol=1e-5, atol=1e-5)
    assert np.allclose(r_o.astype(np.float32), r_h, rtol=1e-5, atol=1e-3)
    ib, tb, bb, mr = compare_states(o.layout, o.get_state(), h.get_state())
    assert ib == 0 and tb == 0, (ib, tb, bb, mr)
    n_oob = 0
    if env_id.endswith("v2"):
        n_oob = len(state_cases.oob_states(env_id, n_agents=n_agents)[0])
        assert d_h[-n_oob:].all() and not t_h[-n_oob:].any()          # every constructed OOB state terminates
        scale = DECAY[1] ** (-DECAY[0]) if decay else 1.0
        assert (r_h[-n_oob:] < -80.0 * scale).all()
    h.close()
    o.close()
    return dict(envs=N, oob=n_oob, dones=int(d_h.sum()), bit_bad=bb)


@pytest.mark.parametrize("env_id", IDS)
@pytest.mark.parametrize("decay", [False, True])
def test_kernel_source_one_step_from_identical_states(env_id, decay):
    from emu_lib import emu_lib
    if decay and not env_id.endswith("v2"):
        pytest.skip("decay applies to v2 only")
    _abi_vs_oracle(env_id, lib=emu_lib(), decay=decay)


def test_kernel_source_v2_out_of_bounds_more_agents():
    from emu_lib import emu_lib
    _abi_vs_oracle("MultiRobotPuzzle-v2", lib=emu_lib(), n_agents=4, decay=True)


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", IDS)
@pytest.mark.parametrize("device_path", [False, True])
@pytest.mark.parametrize("decay", [False, True])
def test_gpu_one_step_from_identical_states(env_id, device_path, decay):
    """CUDA vs oracle incl. the v2 out-of-bounds branches (mrp02:552-563) with and without decay, both entry points."""
    if decay and not env_id.endswith("v2"):
        pytest.skip("decay applies to v2 only")
    rep = _abi_vs_oracle(env_id, device_path=device_path, decay=decay)
    print(env_id, "device" if device_path else "host", rep)


@pytest.mark.gpu
@pytest.mark.parametrize("device_path", [False, True])
def test_gpu_v2_out_of_bounds_more_agents(device_path):
    rep = _abi_vs_oracle("MultiRobotPuzzleHeavy-v2", device_path=device_path, n_agents=5, decay=True)
    print(rep)
