"""Known-answer tests pinning the CPU oracle (oracle/) — hand-derivable cases only.

The reference ships no golden vectors for this path ("parity unpinned"),
so these KATs are what anchors the restatement: Random123's published Philox vectors, the
closed-form mass/inertia of the shapes in reference mrp00:303-332,62-67 / mrp02:331-341,64-67,
Box2D's Pade damping, a face-face box manifold, a head-on wall impulse and an analytic TOI.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle_lib import OracleBatch, StateView, lib, _p


def philox(c, k):
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(C.c_uint32(c[0]), C.c_uint32(c[1]), C.c_uint32(c[2]), C.c_uint32(c[3]), C.c_uint32(k[0]), C.c_uint32(k[1]), _p(out))
    return [int(x) for x in out]


def test_philox_random123_kat():
    # Random123 kat_vectors: philox4x32 10 rounds
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("variant,m,I,cy", [
    (0, 20.0, 17.083333, 0.25),      # rho=5: areas 1 + 3
    (1, 160.0, 546.66667, 0.5),      # rho=10, 2x size
    (2, 0.2496, 0.008528, 0.05),     # rho=1.56
    (3, 3.2, 0.1093333, 0.05),       # rho=20
])
def test_block_mass_closed_form(variant, m, I, cy):
    b = OracleBatch(variant, 1)
    mass, inv_mass, inertia, inv_i, lcx, lcy = b.body_mass(0)
    assert mass == pytest.approx(m, rel=1e-6)
    assert inertia == pytest.approx(I, rel=1e-5)
    assert inv_mass == pytest.approx(1 / m, rel=1e-6)
    assert lcx == 0.0 and lcy == pytest.approx(cy, rel=1e-6)


def test_robot_mass():
    # v0: density 0 => mass 1, invI 0 (b2Body ctor defaults); v2: octagon rho 17.3
    assert list(OracleBatch(0, 1).body_mass(1)[:4]) == [1.0, 1.0, 0.0, 0.0]
    m = OracleBatch(2, 1).body_mass(1)
    # octagon (+-0.039,+-0.095),(+-0.095,+-0.039): area = 0.19^2 - 2*0.056^2
    area = 0.19 ** 2 - 2 * 0.056 ** 2
    assert m[0] == pytest.approx(17.3 * area, rel=1e-5)
    assert m[2] == pytest.approx(0.00245530, rel=1e-4)


def test_octagon_hull_order_and_normals():
    b = OracleBatch(0, 1)
    n, v, nrm, fr = b.fixture(2)
    assert n == 8 and fr == pytest.approx(0.2)
    # b2PolygonShape::Set starts from the right-most vertex, lowest y on ties, CCW
    exp = [(0.75, -0.25), (0.75, 0.25), (0.25, 0.75), (-0.25, 0.75), (-0.75, 0.25), (-0.75, -0.25), (-0.25, -0.75), (0.25, -0.75)]
    assert np.allclose(v, exp)
    assert np.allclose(nrm[0], (1, 0)) and np.allclose(nrm[1], (math.sqrt(.5), math.sqrt(.5)))
    # T-block: stem then bar, wall order L R B T
    assert np.allclose(b.fixture(0)[1], [(-.5, -1), (.5, -1), (.5, 0), (-.5, 0)])
    assert np.allclose(b.fixture(1)[1], [(-1.5, 0), (1.5, 0), (1.5, 1), (-1.5, 1)])
    assert b.fixture(0)[3] == pytest.approx(0.999)


def _free_state(b, variant_n, block=(15.0, 11.0, 0.0), agents=None, vel=None):
    """A contact-free state: block mid-arena, agents spread out."""
    l = b.layout
    w = np.zeros((1, l.state_words), dtype=np.uint32)
    w[0, 1] = 0
    f = w[0, l.off_bodies:l.off_bodies + 6 * l.n_dyn_bodies].view(np.float32).reshape(-1, 6)
    # body c is the centre of mass; fine for these tests
    f[0, :3] = block
    for i in range(l.n_agents):
        f[1 + i, :3] = agents[i] if agents else (3.0 + 2.5 * i, 3.0, 0.0)
    if vel is not None:
        f[:, 3:6] = vel
    aabb = w[0, l.off_aabb:l.off_aabb + 4 * l.n_dyn_fixtures].view(np.float32).reshape(-1, 4)
    # degenerate fat AABBs force a re-insert (and pair discovery) on the first step
    aabb[:] = (1e6, 1e6, 1e6, 1e6)
    d = w[0, l.off_dists:l.off_dists + 2 * (l.n_agents + 1)].view(np.float64)
    d[:] = 100.0
    g = w[0, l.off_goal:l.off_goal + 4].view(np.float64)
    g[:] = (320.0, 262.5)
    return w


def test_damping_pade_and_holonomic_control():
    # v0: v is overwritten by the action each step, then damped by 1/(1+h*5) (A.8) before x += h v
    b = OracleBatch(0, 1)
    b.set_auto_reset(False)
    b.set_state(_free_state(b, 2))
    a = np.array([[1.0, 0.0, 0.5, 0.0, -1.0, 0.0]], dtype=np.float32)
    s0 = StateView(b.layout, b.get_state()).bodies.copy()
    b.step(a)
    s1 = StateView(b.layout, b.get_state()).bodies
    speed = np.float32(10 / 30.0 * 4)
    k = np.float32(1.0) / (np.float32(1.0) + np.float32(0.02) * np.float32(5.0))
    v = speed * k
    assert s1[0, 1, 3] == v and s1[0, 1, 4] == 0.0
    assert s1[0, 1, 0] == np.float32(s0[0, 1, 0] + np.float32(0.02) * v)
    assert s1[0, 1, 5] == np.float32(0.5) * k
    assert s1[0, 2, 4] == -v
    # block: only the (tiny) soft force acts
    assert abs(s1[0, 0, 3]) < 1e-3


def _collide(va, xa, vb, xb):
    va = np.asarray(va, dtype=np.float32); vb = np.asarray(vb, dtype=np.float32)
    out = np.zeros(12, dtype=np.float32)
    lib().orc_collide(len(va), _p(va), _p(np.asarray(xa, dtype=np.float32)), len(vb), _p(vb), _p(np.asarray(xb, dtype=np.float32)), _p(out))
    return out


BOX = [(-1, -1), (1, -1), (1, 1), (-1, 1)]


def test_box_box_face_manifold():
    # two 2x2 boxes, B shifted right by 1.99 (penetration 0.01 < skin): reference face A edge 1 (normal +x)
    o = _collide(BOX, (0, 0, 0), BOX, (1.99, 0.5, 0))
    assert o[0] == 2 and o[1] == 0
    assert tuple(o[2:4]) == (1.0, 0.0) and tuple(o[4:6]) == (1.0, 0.0)
    # incident edge of B is its edge 3 (normal -x): vertices 3 (-1,1) and 0 (-1,-1).  The second clip
    # (A's side plane y = +1 + totalRadius) keeps vertex 0 first, then emits the clipped point of
    # vertex 3 at world y = 1.02, i.e. local y = 1.02 - 0.5 = 0.52 (b2ClipSegmentToLine output order).
    assert tuple(o[6:8]) == (-1.0, -1.0)
    assert o[9] == -1.0 and o[10] == pytest.approx(0.52, abs=1e-6)
    # separated beyond the 0.02 skin => no points
    assert _collide(BOX, (0, 0, 0), BOX, (2.021, 0, 0))[0] == 0
    assert _collide(BOX, (0, 0, 0), BOX, (2.019, 0, 0))[0] == 2


def test_box_box_faceB_choice_and_flip():
    # rotate nothing, but make B's face the deeper-separating axis: A vertex into B's face
    o = _collide(BOX, (0, 0, math.pi / 4), BOX, (2.4, 0, 0))
    # diamond A's corner (sqrt2,0) penetrates B's left face at x=1.4: faceB reference
    assert o[0] == 1 and o[1] == 1
    assert tuple(o[2:4]) == (-1.0, 0.0)


def test_head_on_wall_impulse_and_contact_flags():
    # v0 robot touching the left wall (inner face x=1), pushing into it: normal velocity must be removed,
    # tangential motion survives up to friction 0.2.
    b = OracleBatch(0, 1)
    b.set_auto_reset(False)
    w = _free_state(b, 2, agents=[(1.0 + 0.75 + 0.015, 8.0, 0.0), (12.0, 3.0, 0.0)])
    b.set_state(w)
    a = np.array([[-1.0, 0.0, 0.0, 0.0, 0.0, 0.0]], dtype=np.float32)
    b.step(a)   # contact created at the end of this step (not yet touching): robot moves into the skin
    b.step(a)   # manifold evaluated, solver acts
    sv = StateView(b.layout, b.get_state())
    tab = sv.contact_table(0)
    wall = [c for c in tab if c["fA"] == 2 and c["fB"] == 4]
    assert len(wall) == 1 and wall[0]["touching"] == 1 and wall[0]["pointCount"] == 2
    assert sv.bodies[0, 1, 3] > -1e-6          # no velocity into the wall
    assert sv.bodies[0, 1, 0] > 1.0 + 0.75 - 0.01
    for _ in range(20):
        b.step(a)
    sv = StateView(b.layout, b.get_state())
    # rests within slop of the wall face
    assert 1.75 - 0.02 < sv.bodies[0, 1, 0] < 1.75 + 0.021


def test_toi_box_vs_wall_analytic():
    # 2x2 box moving +x by 1.0 towards a static box whose face is 0.5 away: TOI target separation is
    # linearSlop (0.005) between the core polygons => t = (0.5 - 0.005) / 1.0 within tolerance 0.00125
    t = C.c_float(0)
    va = np.asarray(BOX, dtype=np.float32)
    sA = np.asarray((0, 0, 0, 1.0, 0, 0), dtype=np.float32)
    sB = np.asarray((2.5, 0, 0, 2.5, 0, 0), dtype=np.float32)
    st = lib().orc_toi(4, _p(va), _p(sA), 4, _p(va), _p(sB), C.byref(t))
    assert st == 3  # touching
    assert t.value == pytest.approx(0.495, abs=0.00125 + 1e-6)
    # not reaching: separated
    sA2 = np.asarray((0, 0, 0, 0.4, 0, 0), dtype=np.float32)
    assert lib().orc_toi(4, _p(va), _p(sA2), 4, _p(va), _p(sB), C.byref(t)) == 4 and t.value == 1.0
    # already overlapping cores: overlapped, t = 0
    sA3 = np.asarray((0.6, 0, 0, 1.0, 0, 0), dtype=np.float32)
    assert lib().orc_toi(4, _p(va), _p(sA3), 4, _p(va), _p(sB), C.byref(t)) == 2


def test_reset_is_deterministic_and_sharding_invariant():
    a = OracleBatch(1, 6, seed=17)
    b = OracleBatch(1, 3, seed=17, env_id_base=3)
    oa, ob = a.reset(), b.reset()
    assert np.array_equal(oa[3:], ob)
    assert not np.array_equal(oa[0], oa[1])
    # spawn ranges (mrp00:311-315): block centre of mass stays within the arena
    sv = StateView(a.layout, a.get_state())
    assert (sv.bodies[:, :, 0] > 0.5).all() and (sv.bodies[:, :, 0] < 20.9).all()


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_state_roundtrip_continues_identically(variant):
    a = OracleBatch(variant, 4, seed=5)
    a.reset()
    for t in range(30):
        a.step(a.sample_actions(t))
    w = a.get_state()
    b = OracleBatch(variant, 4, seed=5)
    b.set_state(w)
    assert np.array_equal(b.get_state(), w)
    for t in range(30, 60):
        act = a.sample_actions(t)
        ra, rb = a.step(act), b.step(act)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
    assert np.array_equal(a.get_state(), b.get_state())


def test_v0_reward_and_obs_layout():
    b = OracleBatch(0, 1)
    b.set_auto_reset(False)
    b.set_state(_free_state(b, 2))
    obs, rew, done, trunc = b.step(np.zeros((1, 6), dtype=np.float32))
    sv = StateView(b.layout, b.get_state())
    bod = sv.bodies[0]
    # agent 0 relative position in px (mrp00:447-451)
    assert obs[0, 0] == pytest.approx((bod[1, 0] - bod[0, 0]) * 30, rel=1e-6)
    assert obs[0, 2] == pytest.approx(30 * math.hypot(bod[1, 0] - bod[0, 0], bod[1, 1] - bod[0, 1]), rel=1e-6)
    assert obs[0, 3] == 0.0
    # block rel to goal (320, 262.5) px
    assert obs[0, 8] == pytest.approx(bod[0, 0] * 30 - 320, rel=1e-6)
    assert obs[0, 10] == 0.0  # -(angle mod 2pi), angle 0
    # bar vertex 0 = (-1.5, 0) local, block origin = c - R*(0,0.25)
    assert obs[0, 12] == pytest.approx((bod[0, 0] - 1.5) * 30, rel=1e-5)
    assert obs[0, 13] == pytest.approx((bod[0, 1] - 0.25) * 30, rel=1e-5)
    # reward: prev dists were 100 px (set in the state), weights mrp00:231-239, factor DS/4
    d_blk = obs[0, 11]
    d_ag = [obs[0, 2], obs[0, 6]]
    exp = (100 - d_blk) * 50 / 4 - 0.025 * d_blk / 4 + sum((100 - d) * 10 / 4 - 0.1 * d / 4 for d in d_ag)
    assert rew[0] == pytest.approx(exp, rel=1e-9)
    assert done[0] == 0


def test_v0_completion_reward_and_done():
    b = OracleBatch(0, 1)
    b.set_auto_reset(False)
    # block COM at the goal (10.6667, 8.75) m
    b.set_state(_free_state(b, 2, block=(320 / 30.0, 262.5 / 30.0, 0.3)))
    obs, rew, done, trunc = b.step(np.zeros((1, 6), dtype=np.float32))
    assert done[0] == 1 and trunc[0] == 0
    assert rew[0] > 10000   # +10 in-place delta +10000 completion
    assert StateView(b.layout, b.get_state()).w[0, 2] == 1  # blks_in_place persists


def test_time_limit_truncation():
    b = OracleBatch(0, 1)
    b.set_auto_reset(False)
    w = _free_state(b, 2)
    w[0, 0] = 1999
    b.set_state(w)
    obs, rew, done, trunc = b.step(np.zeros((1, 6), dtype=np.float32))
    assert done[0] == 1 and trunc[0] == 1


def test_v2_out_of_bounds_penalty():
    b = OracleBatch(2, 1)
    b.set_auto_reset(False)
    l = b.layout
    w = np.zeros((1, l.state_words), dtype=np.uint32)
    f = w[0, l.off_bodies:l.off_bodies + 18].view(np.float32).reshape(3, 6)
    f[0, :3] = (1.2857, 0.7232, 0.0)
    f[1, :3] = (0.05, 0.7, 4.712389)      # agent COM inside the 0.1 band -> OOB (mrp02:288-295)
    f[2, :3] = (0.5, 0.4, 4.712389)
    w[0, l.off_aabb:l.off_aabb + 4 * l.n_dyn_fixtures].view(np.float32)[:] = 1e6
    w[0, l.off_dists:l.off_dists + 6].view(np.float64)[:] = 0.5
    w[0, l.off_goal:l.off_goal + 4].view(np.float64)[:] = (0.85, 0.3)
    b.set_state(w)
    obs, rew, done, trunc = b.step(np.zeros((1, 4), dtype=np.float32))
    assert done[0] == 1 and rew[0] < -900
    assert obs.shape[1] == 39 and obs[0, -1] == pytest.approx(0.1)
