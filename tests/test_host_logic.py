"""Host-side mirror of the reference interface: registry ids, spaces, gym-style single env, knobs, sharding."""
import numpy as np
import pytest

import gym_puzzles_b200 as gp
from emu_lib import emu_lib
from gym_puzzles_b200 import spaces
from gym_puzzles_b200.vector_env import shard_range


def test_registry_matches_reference_registration():
    # reference gym_puzzles/__init__.py:3-29; MultiRobotPuzzleSquare-v2 is this package's extension (BASELINE.json configs[4])
    assert {k: (v.max_episode_steps, v.reward_threshold) for k, v in gp.registry.items() if k != "MultiRobotPuzzleSquare-v2"} == {
        "MultiRobotPuzzle-v0": (2000, 500), "MultiRobotPuzzleHeavy-v0": (3000, 500),
        "MultiRobotPuzzle-v2": (2000, 500), "MultiRobotPuzzleHeavy-v2": (2000, 500)}
    with pytest.raises(KeyError):
        gp.spec("MultiRobotPuzzle-v3")


@pytest.mark.parametrize("env_id,n,obs,act", [("MultiRobotPuzzle-v0", 2, 28, 6), ("MultiRobotPuzzleHeavy-v0", 5, 40, 15),
                                              ("MultiRobotPuzzle-v2", 2, 39, 4), ("MultiRobotPuzzleHeavy-v2", 2, 39, 4)])
def test_spaces(env_id, n, obs, act):
    o, a = spaces.observation_space(env_id, n), spaces.action_space(env_id, n)
    assert o.shape == (obs,) and a.shape == (act,)
    assert o.dtype == np.float32 and a.dtype == np.float32
    assert np.all(a.high == 1) and np.all(a.low == -1)
    # theta threshold 2*pi on the block angle slot (reference mrp00:186-193, mrp02:175-186)
    idx = 4 * n + 2 if env_id.endswith("v0") else 9 * n + 2
    assert o.high[idx] == pytest.approx(2 * np.pi)
    assert np.isinf(o.high[0])


def test_gym_style_env_surface():
    env = gp.make("MultiRobotPuzzle-v0", _lib=emu_lib())
    assert env.spec.max_episode_steps == 2000
    obs = env.reset()
    assert obs.shape == (28,) and obs.dtype == np.float32
    o2, r, d, info = env.step(env.action_space.sample())
    assert o2.shape == (28,) and isinstance(r, float) and isinstance(d, bool) and info == {}
    assert env.get_deltaAgent() == 10 and env.get_agentDist() == 0.1 and env.get_deltaBlk() == 50 and env.get_blkDist() == 0.025
    env.set_reward_params(agentDelta=1, agentDistance=2, blockDelta=3, blockDistance=4)
    assert (env.get_deltaAgent(), env.get_agentDist(), env.get_deltaBlk(), env.get_blkDist()) == (1, 2, 3, 4)
    with pytest.raises(NotImplementedError):
        env.render()
    with pytest.raises(ValueError):
        env.step(np.zeros(5))
    assert env._return_status() == "Stayed in bounds"


def test_seed_makes_resets_reproducible():
    a, b = gp.make("MultiRobotPuzzleHeavy-v0", _lib=emu_lib()), gp.make("MultiRobotPuzzleHeavy-v0", _lib=emu_lib())
    a.seed(123); b.seed(123)
    assert np.array_equal(a.reset(), b.reset())
    assert not np.array_equal(a.reset(), a.reset())   # next episode differs
    b.seed(124)
    assert not np.array_equal(a.seed(123) and a.reset(), b.reset())


def test_v2_knobs():
    env = gp.make("MultiRobotPuzzle-v2", _lib=emu_lib())
    assert env.num_agents == 2 and env.observation_space.shape == (39,)
    obs = env.reset()
    assert obs[-1] == pytest.approx(0.1)          # scaled_epsilon appended (reference mrp02:531-532)
    env.update_goal(epoch=5, nb_epochs=10)        # EPSILON * (2 - 0.5)
    assert env.step(np.zeros(4))[0][-1] == pytest.approx(0.15)
    env.update_params(timestep=2, decay=0.5)      # decay**(-t) = 4
    assert env._h.get_params()["decay_pow"] == pytest.approx(4.0)
    assert (env.get_deltaAgent(), env.get_agentDist(), env.get_deltaBlk(), env.get_blkDist()) == (10, 0.25, 25, 0.1)


def test_time_limit_in_single_env():
    env = gp.make("MultiRobotPuzzle-v0", _lib=emu_lib())
    env.reset()
    w = env.get_state().copy()
    w[0] = 1999
    env.set_state(w)
    _, _, done, info = env.step(np.zeros(6))
    assert done and info.get("TimeLimit.truncated") is True


def test_shard_range():
    assert shard_range(1048576, 3, 8) == (393216, 131072)
    with pytest.raises(ValueError):
        shard_range(10, 0, 3)
