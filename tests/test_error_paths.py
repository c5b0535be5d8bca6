"""Error behaviour of the C-ABI (include/mrp_b200.h: "no exceptions across the ABI", status < 0 + mrp_last_error), checked
through the host build of the same source.  The reference raises Python exceptions / Box2D asserts in these situations
("Ownership / errors")."""
import ctypes as C

import numpy as np
import pytest
from emu_lib import emu_lib

from gym_puzzles_b200 import abi


@pytest.mark.parametrize("n", [0, -3])
def test_create_rejects_empty_batches(n):
    with pytest.raises(abi.MrpError, match="num_envs must be > 0"):
        abi.Handle("MultiRobotPuzzle-v0", n, seed=1, lib=emu_lib())


def test_create_rejects_unsupported_robot_counts():
    with pytest.raises(abi.MrpError, match="n_agents"):
        abi.Handle("MultiRobotPuzzle-v2", 4, seed=1, lib=emu_lib(), n_agents=9)


def test_create_rejects_unknown_variant_and_null_arguments():
    lib = emu_lib()
    cfg = abi.Config()
    cfg.variant, cfg.num_envs = 17, 4
    out = C.c_void_p()
    assert lib.lib.mrp_create(C.byref(cfg), C.byref(out)) < 0 and not out.value
    assert b"variant" in lib.lib.mrp_last_error()
    assert lib.lib.mrp_create(None, C.byref(out)) < 0
    assert lib.lib.mrp_step(None, None, None) < 0 and b"null" in lib.lib.mrp_last_error()
    assert lib.lib.mrp_step_host(None, None, None, None, None, None) < 0


def test_step_host_checks_the_action_shape_and_accepts_missing_outputs():
    h = abi.Handle("MultiRobotPuzzle-v0", 3, seed=1, lib=emu_lib())
    h.reset_host()
    with pytest.raises(ValueError):
        h.step_host(np.zeros((2, 6), np.float32))
    # any output pointer may be NULL (header): only the action pointer is required
    a = np.zeros((3, 6), np.float32)
    assert h.lib.lib.mrp_step_host(h.h, a.ctypes.data_as(C.c_void_p), None, None, None, None) == 0
    assert h.lib.lib.mrp_step_host(h.h, None, None, None, None, None) < 0
    h.close()


def test_actions_outside_the_box_are_not_clipped():
    """the reference does not clip actions: a larger action gives a larger commanded velocity"""
    h1 = abi.Handle("MultiRobotPuzzle-v0", 1, seed=3, lib=emu_lib())
    h2 = abi.Handle("MultiRobotPuzzle-v0", 1, seed=3, lib=emu_lib())
    assert np.array_equal(h1.reset_host(), h2.reset_host())
    a = np.zeros((1, 6), np.float32)
    a[0, 0] = 1.0
    o1 = h1.step_host(a)[0]
    o2 = h2.step_host(3.0 * a)[0]
    assert not np.array_equal(o1, o2) and np.isfinite(o2).all()
    h1.close(); h2.close()


def _check_nan_guard(lib):
    """a body whose state became NaN / inf ends its episode by force (done = trunc = 1, reward 0),
    is respawned by the auto-reset and counted in MRP_STAT_NAN_RESETS; the other envs are untouched."""
    kw = {} if lib is None else {"lib": lib}
    N = 40
    ref = abi.Handle("MultiRobotPuzzleHeavy-v0", N, seed=9, **kw)
    h = abi.Handle("MultiRobotPuzzleHeavy-v0", N, seed=9, **kw)
    ref.reset_host(); h.reset_host()
    w = h.get_state()
    l = h.layout
    bad = [3, 17]
    for i, e in enumerate(bad):
        f = w[e, l.off_bodies:l.off_bodies + 6 * l.n_dyn_bodies].view(np.float32)
        if i == 0:
            f[6 * 1 + 0] = np.nan          # a robot's position (its velocity would be overwritten by the holonomic control)
        else:
            f[6 * 0 + 3] = np.inf          # the block's velocity
    h.set_state(w)
    a = np.random.default_rng(0).uniform(-1, 1, (N, h.act_dim)).astype(np.float32)
    obs_r, rew_r, done_r, tr_r = ref.step_host(a)
    obs, rew, done, trunc = h.step_host(a)
    assert done[bad].all() and trunc[bad].all() and (rew[bad] == 0).all()
    assert np.isfinite(obs).all()                               # the rows hold the respawned envs' observations
    good = np.setdiff1d(np.arange(N), bad)
    assert np.array_equal(obs[good], obs_r[good]) and np.array_equal(done[good], done_r[good])
    s = h.stats()
    assert s["nan_resets"] == 2 and s["episodes"] >= 2 and np.isfinite(s["sum_return"])
    assert np.isfinite(h.get_state()[:, l.off_bodies:l.off_bodies + 6 * l.n_dyn_bodies].view(np.float32)).all()
    h.close(); ref.close()


def test_nan_guard_kernel_source():
    from emu_lib import emu_lib
    _check_nan_guard(emu_lib())


@pytest.mark.gpu
def test_nan_guard_gpu():
    _check_nan_guard(None)
