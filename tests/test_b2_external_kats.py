"""Known-answer tests whose answers do NOT come from this repository's Box2D restatement:
the 2x2 LCP of Box2D's block solver solved by hand for all four of its cases, a body pressed against a wall by a constant
force (impulse = force x dt, shared equally by the two manifold points), and a time of impact of a ROTATING body checked
against closed-form geometry.  The world is built through the pybox2d stand-in (tests/refshim) over oracle/b2core.hpp."""
import ctypes as C
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim", "standins"))
import Box2D  # noqa: E402  (the stand-in; nothing else in the process sees it under this name)
from oracle_lib import _p, lib  # noqa: E402

sys.path.pop(0)
sys.modules.pop("Box2D", None)

A = 0.5          # half extent of the dynamic box
SEP = 0.01       # gap between the core polygons (inside the 2 * polygonRadius = 0.02 skin: a manifold exists)


def _world(vy, omega, friction=0.0):
    w = Box2D.b2World(gravity=(0, 0), doSleep=False)
    wall = w.CreateStaticBody(position=(0.0, -0.5), fixtures=Box2D.fixtureDef(shape=Box2D.polygonShape(box=(5.0, 0.5)), friction=friction))
    box = w.CreateDynamicBody(position=(0.0, A + SEP), angle=0.0, linearDamping=0.0, angularDamping=0.0,
                              fixtures=Box2D.fixtureDef(shape=Box2D.polygonShape(box=(A, A)), density=1.0, friction=friction))
    box.linearVelocity = (0.0, vy)
    box.angularVelocity = float(omega)
    return w, wall, box


def _lcp(vy, omega, m, inertia):
    """Box2D block solver (b2ContactSolver::SolveVelocityConstraints, 'Block solver' note): vn = K x + b, vn >= 0, x >= 0,
    vn_i x_i = 0 for the two points of a manifold, solved by enumerating the four cases — in float64, by hand."""
    rx = np.array([-A, A])                               # contact points midway between the faces: r = (+-a, -(a + sep / 2))
    b = vy + omega * rx                                  # normal velocity of each point, n = +y
    K = 1.0 / m + np.outer(rx, rx) / inertia             # (r_i x n)(r_j x n) = r_ix r_jx
    x = -np.linalg.solve(K, b)
    if (x >= 0).all():
        return x, 1
    x1 = -b[0] / K[0, 0]
    if x1 >= 0 and K[0, 1] * x1 + b[1] >= 0:
        return np.array([x1, 0.0]), 2
    x2 = -b[1] / K[1, 1]
    if x2 >= 0 and K[0, 1] * x2 + b[0] >= 0:
        return np.array([0.0, x2]), 3
    assert (b >= 0).all()
    return np.zeros(2), 4


@pytest.mark.parametrize("vy,omega,case", [(-1.0, 0.0, 1), (-1.0, 4.0, 2), (-1.0, -4.0, 3), (1.0, 0.0, 4), (-0.3, 1.0, 2), (-1.0, 0.5, 1), (-2.0, -9.0, 3)])
def test_block_solver_matches_hand_solved_lcp(vy, omega, case):
    w, wall, box = _world(vy, omega)
    m, inertia = box.mass, box.inertia
    assert m == pytest.approx(1.0, rel=1e-6) and inertia == pytest.approx(2.0 * A * A / 3.0, rel=1e-6)   # unit square, density 1
    x, got_case = _lcp(vy, omega, m, inertia)
    assert got_case == case
    w.Step(1.0 / 50, 180, 60)
    (c,) = w.contacts
    assert c["touching"] and c["pointCount"] == 2
    # map the manifold points to (left, right) by the x coordinate of their local point (in the box's frame)
    pts = sorted(c["points"], key=lambda p: p["localPoint"][0])
    got = np.array([pts[0]["normalImpulse"], pts[1]["normalImpulse"]])
    assert np.allclose(got, x, rtol=2e-5, atol=2e-6), (got, x)
    assert all(abs(p["tangentImpulse"]) < 1e-7 for p in pts)
    # velocities after the solve: v' = v + sum(x) / m, w' = w + sum(r_x x) / I
    v_after = vy + x.sum() / m
    w_after = omega + (-A * x[0] + A * x[1]) / inertia
    assert box.linearVelocity[1] == pytest.approx(v_after, abs=5e-6)
    assert box.angularVelocity == pytest.approx(w_after, abs=5e-5)


def test_constant_force_is_balanced_by_impulse_force_times_dt():
    """A box pushed against the wall by F for many steps comes to rest; per step the contact returns exactly the momentum
    the force adds: sum of the normal impulses = F * dt, half on each point (symmetry); warm starting keeps it there."""
    w, wall, box = _world(0.0, 0.0)
    F, dt = 10.0, 1.0 / 50
    for _ in range(200):
        box.ApplyForce((0.0, -F), box.worldCenter, True)
        w.Step(dt, 180, 60)
    (c,) = w.contacts
    imp = [p["normalImpulse"] for p in c["points"]]
    assert len(imp) == 2 and sum(imp) == pytest.approx(F * dt, rel=1e-4)
    assert imp[0] == pytest.approx(imp[1], rel=1e-3)
    assert abs(box.linearVelocity[1]) < 1e-4 and abs(box.angularVelocity) < 1e-4
    # rests inside the skin: core faces between 0 and 2 * polygonRadius apart, pushed out towards -linearSlop penetration
    gap = box.position[1] - A
    assert -0.005 - 1e-4 < gap - 0.02 < 1e-4


def test_toi_of_a_rotating_rod_matches_closed_form_geometry():
    """b2TimeOfImpact for a rod swinging down onto a wall: at the returned time the distance between the core polygons —
    here simply the height of the rod's lower right corner above the wall's top face — equals the target separation
    (linearSlop = 0.005) within the tolerance 0.25 * linearSlop."""
    wall = np.array([[-5, -1], [5, -1], [5, 1], [-5, 1]], dtype=np.float32)       # centred at (0, -1.5): top face y = -0.5
    rod = np.array([[-1, -0.05], [1, -0.05], [1, 0.05], [-1, 0.05]], dtype=np.float32)
    a1 = -1.0                                                                       # swings clockwise by 1 rad about its centre
    sW = np.asarray((0, -1.5, 0, 0, -1.5, 0), dtype=np.float32)
    sR = np.asarray((0, 0, 0, 0, 0, a1), dtype=np.float32)
    t = C.c_float(0)
    st = lib().orc_toi(4, _p(wall), _p(sW), 4, _p(rod), _p(sR), C.byref(t))
    assert st == 3   # e_touching

    def height(tt):                      # lower right corner (1, -0.05) of the rod above y = -0.5
        th = a1 * tt
        return math.sin(th) * 1.0 + math.cos(th) * (-0.05) + 0.5

    assert abs(height(t.value) - 0.005) <= 0.00125 + 1e-6
    # closed form: sin(th) - 0.05 cos(th) = -0.495  =>  th = asin(-0.495 / R) + atan2(0.05, 1), R = sqrt(1 + 0.05^2)
    R = math.hypot(1.0, 0.05)
    th = math.asin(-0.495 / R) + math.atan2(0.05, 1.0)
    assert t.value == pytest.approx(th / a1, abs=0.003)
    # a swing that stops short of the wall: separated, t = 1
    sR2 = np.asarray((0, 0, 0, 0, 0, -0.3), dtype=np.float32)
    assert lib().orc_toi(4, _p(wall), _p(sW), 4, _p(rod), _p(sR2), C.byref(t)) == 4 and t.value == 1.0


@pytest.mark.parametrize("fork_230", [0, 1])
def test_hello_box2d_listing_of_the_box2d_manual(fork_230):
    """The one worked example Box2D itself publishes with numbers: "Hello Box2D" (Box2D v2.x manual, chapter 2; the program
    is HelloWorld.cpp of the distribution).  A static ground box (half extents 50 x 10 at (0, -10)), a dynamic unit box
    (half extents 1 x 1, density 1, friction 0.3) dropped from (0, 4) under gravity (0, -10), stepped 60 times with
    timeStep = 1/60, 6 velocity and 2 position iterations, printing position and angle with "%4.2f %4.2f %4.2f".  The manual
    lists the output as
        0.00 4.00 0.00 / 0.00 3.99 0.00 / 0.00 3.98 0.00 / ... / 0.00 1.25 0.00 / 0.00 1.13 0.00 / 0.00 1.01 0.00
    i.e. free fall through step 45 (1.13 is 1.1250004 in float32: the accumulated rounding decides the digit) and, in step 46,
    the continuous-collision path: the fall would end at 0.997, b2TimeOfImpact stops the box linearSlop above the ground's
    skin and the TOI sub-step leaves it at 1.0146 — "1.01", where it stays.  Exercises the integrator, fat-AABB pair creation,
    b2TimeOfImpact / SolveTOI against a static body, the contact solver and the position correction on numbers that come from
    Box2D's own documentation.  The oracle's world has no gravity term (the reference uses gravity (0, 0)); gravity is applied
    as the force m g, which is the same arithmetic: v += h * (invMass * force) with invMass * (m g) = g exactly for m = 4."""
    # both collision forks (Box2D 2.3.0 / >= 2.3.1, oracle/b2core.hpp g_fork_230) have to print the same listing: the manual's output did not
    # change between those releases
    shim = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "libb2shim.so"))
    old = shim.b2s_set_box2d_fork(fork_230)
    try:
        _hello_box2d()
    finally:
        shim.b2s_set_box2d_fork(old)


def _hello_box2d():
    w = Box2D.b2World(gravity=(0, 0), doSleep=False)
    w.CreateStaticBody(position=(0.0, -10.0), fixtures=Box2D.fixtureDef(shape=Box2D.polygonShape(box=(50.0, 10.0))))
    box = w.CreateDynamicBody(position=(0.0, 4.0), angle=0.0, linearDamping=0.0, angularDamping=0.0,
                              fixtures=Box2D.fixtureDef(shape=Box2D.polygonShape(box=(1.0, 1.0)), density=1.0, friction=0.3))
    assert box.mass == 4.0
    lines = []
    for _ in range(60):
        box.ApplyForce((0.0, -10.0 * box.mass), box.worldCenter, True)
        w.Step(1.0 / 60.0, 6, 2)
        lines.append("%4.2f %4.2f %4.2f" % (box.position[0], box.position[1], box.angle))
    assert lines[:3] == ["0.00 4.00 0.00", "0.00 3.99 0.00", "0.00 3.98 0.00"]
    assert lines[43:46] == ["0.00 1.25 0.00", "0.00 1.13 0.00", "0.00 1.01 0.00"]
    assert all(ln == "0.00 1.01 0.00" for ln in lines[45:])           # "the box lands on the ground box and comes to rest"
    # free fall of the semi-implicit Euler integrator: y_n = 4 - n (n + 1) / 720 while nothing touches
    ys = [float(ln.split()[1]) for ln in lines[:45]]
    assert all(abs(y - (4.0 - n * (n + 1) / 720.0)) <= 0.00501 for n, y in enumerate(ys, start=1))
