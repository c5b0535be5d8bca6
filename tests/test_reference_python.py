"""Live cross-check in the build container: the reference's own Python env classes (imported unmodified from
/root/reference, running over the Box2D stand-in of tests/refshim) against the oracle on identical draws.

Skipped where /root/reference does not exist (the GPU box): the committed fixtures of tests/golden/ carry the same
comparison there (tests/test_golden.py)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim"))
import harness  # noqa: E402
from oracle_lib import OracleBatch, StateView  # noqa: E402

pytestmark = pytest.mark.skipif(not harness.available(), reason="reference checkout not present")
IDS = list(harness.REGISTRY)


def _close(a, b, amplify=1.0):   # see tests/test_golden.py::_close64
    return np.allclose(a, b, rtol=4e-15, atol=1e-13 * amplify)


def _pair(env_id, gid, cap, **kw):
    r = harness.ReferenceEnv(env_id, seed=23, gid=gid, max_episode_steps=cap, **kw)
    o = OracleBatch(env_id, 1, seed=23, env_id_base=gid, max_episode_steps=cap, n_agents=kw.get("num_agents", 0) or 0)
    return r, o


def _lockstep(r, o, actions):
    for t, a in enumerate(actions):
        oobs, orew, odone, otr = o.step(a)
        robs, rrew, rdone, rtr = r.step(a)
        assert _close(robs, oobs[0]) and _close(rrew, orew[0], 1e3), t
        assert rdone == bool(odone[0]) and rtr == bool(otr[0]), t
        sv = StateView(o.layout, o.get_state())
        assert r.goal_contacts == [bool(x) for x in sv.goal_contact[0]], t
        assert np.array_equal(r.body_rows().view(np.uint32), sv.bodies[0].view(np.uint32)), t


@pytest.mark.parametrize("env_id", IDS)
def test_random_rollout_matches(env_id):
    for gid in (3, 77):
        r, o = _pair(env_id, gid, cap=35)
        assert _close(r.reset(), o.reset()[0])
        acts = [o.sample_actions(t)[0] for t in range(90)]
        _lockstep(r, o, acts)


def test_spaces_match_reference():
    from gym_puzzles_b200 import spaces
    for env_id in IDS:
        r = harness.ReferenceEnv(env_id)
        ref_obs, ref_act = r.env.observation_space, r.env.action_space
        n = len(r.env.agents)
        obs_space, act_space = spaces.observation_space(env_id, n), spaces.action_space(env_id, n)
        assert obs_space.shape == ref_obs.shape and act_space.shape == ref_act.shape
        assert np.array_equal(obs_space.high, ref_obs.high) and np.array_equal(obs_space.low, ref_obs.low)
        assert np.array_equal(act_space.high, ref_act.high) and np.array_equal(act_space.low, ref_act.low)
        assert obs_space.dtype == ref_obs.dtype and act_space.dtype == ref_act.dtype


@pytest.mark.parametrize("env_id", ["MultiRobotPuzzle-v2", "MultiRobotPuzzleHeavy-v2"])
def test_v2_completion_branch_from_identical_state(env_id):
    """mrp02:565-582: goal moved onto the block in both => done with puzzleComp * (#contact / n) (contact_weight)."""
    r, o = _pair(env_id, 9, cap=100)
    r.reset(); o.reset()
    x, y = r.env.norm_units(r.env.goal_block.worldCenter)
    r.env.block_final_pos = {"t_block": (x + 0.03, y - 0.02, 0)}
    r.env._calculate_distance()                      # prev distance of the next step, as reset() would have left it
    w = o.get_state()
    l = o.layout
    goal = np.array([x + 0.03, y - 0.02], dtype=np.float64)
    w[0, l.off_goal:l.off_goal + 4] = goal.view(np.uint32)
    d = np.ascontiguousarray(w[0, l.off_dists:l.off_dists + 2 * (l.n_agents + 1)]).view(np.float64)
    d[l.n_agents] = r.env.block_distance["t_block"]
    w[0, l.off_dists:l.off_dists + 2 * (l.n_agents + 1)] = d.view(np.uint32)
    o.set_state(w)
    a = o.sample_actions(5)[0]
    oobs, orew, odone, otr = o.step(a)
    obs, rew, done, info = r.env.step(a.astype(np.float64))
    assert done and odone[0] == 1 and otr[0] == 0
    assert _close(rew, orew[0], 1e3)


@pytest.mark.parametrize("env_id", ["MultiRobotPuzzle-v0", "MultiRobotPuzzle-v2"])
def test_reward_knobs_follow_reference(env_id):
    """set_reward_params / update_goal / update_params (mrp00:231-246, mrp02:216-233)."""
    r, o = _pair(env_id, 4, cap=60)
    kw = dict(agentDelta=3.0, agentDistance=0.5, blockDelta=7.0, blockDistance=0.25, puzzleComp=123.0, outOfBounds=9.0, blkOutOfBounds=4.0)
    r.env.set_reward_params(**kw)
    p = o.get_params()
    p[:7] = [kw[k] for k in ("agentDelta", "agentDistance", "blockDelta", "blockDistance", "puzzleComp", "outOfBounds", "blkOutOfBounds")]
    if env_id.endswith("v2"):
        r.env.update_goal(3, 10)                     # scaled_epsilon = EPSILON * (2 - 3/10)
        r.env.update_params(4, 0.9)                  # decay ** (-timestep)
        p[7] = r.env.scaled_epsilon
        p[8] = 0.9 ** (-4)
    o.set_params(p)
    assert _close(r.reset(), o.reset()[0])
    _lockstep(r, o, [o.sample_actions(t)[0] for t in range(50)])


def test_v2_more_agents_matches_reference():
    """num_agents kw of MultiRobotPuzzle2 (mrp02:139): the oracle follows the reference for n = 4 as well."""
    r, o = _pair("MultiRobotPuzzle-v2", 2, cap=40, num_agents=4)
    assert _close(r.reset(), o.reset()[0])
    _lockstep(r, o, [o.sample_actions(t)[0] for t in range(60)])
