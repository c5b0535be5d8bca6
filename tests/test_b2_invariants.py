"""Physics invariants of the oracle's Box2D restatement (oracle/b2core.hpp), driven through the generic world API of
oracle/libb2shim.so.  Box2D's arithmetic cannot be pinned against pybox2d here; these are properties any
faithful sequential-impulse step must satisfy whatever the Box2D minor version:

  * contact impulses are equal and opposite: linear and angular momentum of a wall-free scene are conserved by a step;
  * restitution 0 and Coulomb friction never add kinetic energy;
  * bodies that do not touch move exactly by the damped symplectic-Euler formulas;
  * a penetrating pair is pushed apart towards the slop allowance and never pulled together.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

import oracle_lib  # noqa: F401  (builds oracle/*.so)

_L = C.CDLL(os.path.join(oracle_lib.ORACLE_DIR, "libb2shim.so"))
_f, _i, _p = C.c_float, C.c_int, C.c_void_p
_L.b2s_world_new.restype = _p
_L.b2s_world_free.argtypes = [_p]
_L.b2s_create_body.argtypes = [_p, _i, _f, _f, _f, _f, _f]
_L.b2s_create_box_fixture.argtypes = [_p, _i, _f, _f, _i, _f, _f, _f, _f, _f, _f]
_L.b2s_create_poly_fixture.argtypes = [_p, _i, _p, _i, _f, _f, _f]
_L.b2s_body_get.argtypes = [_p, _i, _p]
_L.b2s_set_linear_velocity.argtypes = [_p, _i, _f, _f]
_L.b2s_set_angular_velocity.argtypes = [_p, _i, _f]
_L.b2s_step.argtypes = [_p, _f, _i, _i, _p, _i]

OCT = np.array([(-0.25, -0.75), (0.25, -0.75), (0.75, -0.25), (0.75, 0.25), (0.25, 0.75), (-0.25, 0.75), (-0.75, 0.25),
                (-0.75, -0.25)], dtype=np.float32)   # reference mrp00:62-67


class World:
    def __init__(self):
        self.h = _L.b2s_world_new()
        self.bodies = []

    def __del__(self):
        _L.b2s_world_free(self.h)

    def body(self, x, y, angle, shape, density, friction, vx=0.0, vy=0.0, w=0.0, damping=0.0):
        b = _L.b2s_create_body(self.h, 1, x, y, angle, damping, damping)
        if shape == "box":
            _L.b2s_create_box_fixture(self.h, b, 0.6, 0.4, 0, 0, 0, 0, density, friction, 0.0)
        else:
            _L.b2s_create_poly_fixture(self.h, b, OCT.ctypes.data, 8, density, friction, 0.0)
        _L.b2s_set_linear_velocity(self.h, b, vx, vy)
        _L.b2s_set_angular_velocity(self.h, b, w)
        self.bodies.append(b)
        return b

    def get(self, b):
        buf = (C.c_float * 12)()
        _L.b2s_body_get(self.h, b, buf)
        g = np.array(buf[:], dtype=np.float64)
        return dict(c=g[3:5], a=g[2], v=g[5:7], w=g[7], m=g[8], I=g[9])   # inertia about the centre (localCenter = 0 here)

    def step(self, n=1):
        ev = (C.c_int * 320)()
        total = 0
        for _ in range(n):
            total += _L.b2s_step(self.h, 1.0 / 50, 180, 60, ev, 64)
        return total

    def momentum(self):
        P, Lz, KE = np.zeros(2), 0.0, 0.0
        for b in self.bodies:
            g = self.get(b)
            P += g["m"] * g["v"]
            Lz += g["m"] * (g["c"][0] * g["v"][1] - g["c"][1] * g["v"][0]) + g["I"] * g["w"]
            KE += 0.5 * g["m"] * g["v"] @ g["v"] + 0.5 * g["I"] * g["w"] ** 2
        return P, Lz, KE


@pytest.mark.parametrize("seed", range(12))
def test_momentum_conserved_and_energy_not_created(seed):
    rng = np.random.default_rng(seed)
    w = World()
    # two or three bodies flying into each other around the origin, no walls, no damping
    k = 2 + seed % 2
    for j in range(k):
        ang = 2 * math.pi * j / k + rng.uniform(-0.2, 0.2)
        r = rng.uniform(0.9, 1.3)
        speed = rng.uniform(1.0, 6.0)
        w.body(r * math.cos(ang), r * math.sin(ang), rng.uniform(0, 6.28), "box" if (seed + j) % 2 else "oct",
               density=rng.uniform(0.5, 5.0), friction=rng.uniform(0.0, 1.0),
               vx=-speed * math.cos(ang) + rng.uniform(-0.5, 0.5), vy=-speed * math.sin(ang) + rng.uniform(-0.5, 0.5),
               w=rng.uniform(-2, 2))
    P0, L0, KE0 = w.momentum()
    events = 0
    for _ in range(40):
        before = [w.get(b) for b in w.bodies]
        events += w.step()
        after = [w.get(b) for b in w.bodies]
        P, _, KE = w.momentum()
        scale = max(1.0, np.abs(P0).max(), abs(L0))
        assert np.allclose(P, P0, atol=2e-4 * scale), (P, P0)
        # every contact impulse acts at ONE world point on both bodies (arms taken from the pre-step centres), so the
        # angular impulse about the origin sums to zero.  (Positions are then corrected without touching velocities,
        # which is why m c x v itself is not an invariant of a Box2D step.)
        dL = sum(g0["m"] * (g0["c"][0] * (g1["v"][1] - g0["v"][1]) - g0["c"][1] * (g1["v"][0] - g0["v"][0])) + g0["I"] * (g1["w"] - g0["w"])
                 for g0, g1 in zip(before, after))
        assert abs(dL) <= 2e-4 * scale, dL
        assert KE <= KE0 * (1 + 1e-5) + 1e-6, (KE, KE0)
        KE0 = KE
    assert events >= 1          # the bodies did collide (BeginContact fired)


def test_free_flight_is_damped_symplectic_euler():
    w = World()
    b = w.body(3.0, 4.0, 0.3, "box", 2.0, 0.5, vx=1.5, vy=-0.75, w=0.9, damping=5.0)
    c, a, v, om = np.float32([3.0, 4.0]), np.float32(0.3), np.float32([1.5, -0.75]), np.float32(0.9)
    h, k = np.float32(1.0 / 50), np.float32(1.0) / (np.float32(1.0) + np.float32(1.0 / 50) * np.float32(5.0))
    for _ in range(25):
        w.step()
        v = v * k                  # v *= 1 / (1 + h * damping)  (Pade form, b2Island::Solve)
        om = np.float32(om * k)
        c = c + h * v              # x += h * v
        a = np.float32(a + h * om)
        g = w.get(b)
        assert np.array_equal(g["c"].astype(np.float32), c) and np.float32(g["a"]) == a
        assert np.array_equal(g["v"].astype(np.float32), v) and np.float32(g["w"]) == om


def test_penetration_is_pushed_out_towards_slop():
    w = World()
    a = w.body(0.0, 0.0, 0.0, "box", 1.0, 0.3)
    b = w.body(1.0, 0.0, 0.0, "box", 1.0, 0.3)          # boxes are 1.2 wide: 0.2 overlap along x
    gaps = []
    for _ in range(12):
        w.step()
        gaps.append(float(w.get(b)["c"][0] - w.get(a)["c"][0]) - 1.2)
    assert all(g2 >= g1 - 1e-6 for g1, g2 in zip(gaps, gaps[1:]))          # never pulled together
    assert gaps[0] > -0.2 + 0.02                                           # first step already corrects
    # the manifold separation is the core-polygon gap minus the skin 2 * polygonRadius = 0.02; the position solver stops
    # at separation >= -3 * b2_linearSlop, i.e. the cores end between 0.005 and 0.02 apart
    assert 0.02 - 3 * 0.005 - 1e-4 <= gaps[-1] <= 0.02 + 1e-4
    ga, gb = w.get(a), w.get(b)
    assert abs(ga["c"][1]) < 2e-3 and abs(gb["c"][1]) < 2e-3               # (sequential point order tilts the pair by < 1 mrad)
    assert abs((ga["c"][0] + gb["c"][0]) - 1.0) < 1e-5                     # equal masses: pushed apart symmetrically
