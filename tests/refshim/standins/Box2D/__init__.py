"""Stand-in for pybox2d (`Box2D`), backed by the oracle's Box2D restatement (oracle/libb2shim.so).

Only what the reference env modules use (SURVEY.md §8c call-site list).  Arithmetic widths follow pybox2d:
b2Vec2 stores float32 and its operators (`v * s`, `s * v`, `-v`, `b2Dot`) compute in float32; attribute reads hand
Python floats (float64 copies of the float32 values); every argument that crosses into Box2D is rounded to float32.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "..", "..", "..", "..", "oracle", "libb2shim.so")
_L = C.CDLL(os.path.abspath(_LIB))
_f, _i, _p = C.c_float, C.c_int, C.c_void_p
_L.b2s_world_new.restype = _p
_L.b2s_world_free.argtypes = [_p]
_L.b2s_world_clear.argtypes = [_p]
_L.b2s_create_body.argtypes = [_p, _i, _f, _f, _f, _f, _f]
_L.b2s_create_box_fixture.argtypes = [_p, _i, _f, _f, _i, _f, _f, _f, _f, _f, _f]
_L.b2s_create_poly_fixture.argtypes = [_p, _i, _p, _i, _f, _f, _f]
_L.b2s_fixture_vertices.argtypes = [_p, _i, _p]
_L.b2s_fixture_body.argtypes = [_p, _i]
_L.b2s_body_fixtures.argtypes = [_p, _i, _p, _i]
_L.b2s_body_get.argtypes = [_p, _i, _p]
_L.b2s_set_linear_velocity.argtypes = [_p, _i, _f, _f]
_L.b2s_set_angular_velocity.argtypes = [_p, _i, _f]
_L.b2s_apply_force.argtypes = [_p, _i, _f, _f, _f, _f]
_L.b2s_apply_torque.argtypes = [_p, _i, _f]
_L.b2s_apply_linear_impulse.argtypes = [_p, _i, _f, _f, _f, _f]
_L.b2s_apply_angular_impulse.argtypes = [_p, _i, _f]
_L.b2s_world_point.argtypes = [_p, _i, _f, _f, _p]
_L.b2s_world_vector.argtypes = [_p, _i, _f, _f, _p]
_L.b2s_step.argtypes = [_p, _f, _i, _i, _p, _i]
_L.b2s_toi_events.argtypes = [_p]
_L.b2s_load_state.argtypes = [_p, _i, _p, _i, _p, _i, _p]
_L.b2s_contact_get.argtypes = [_p, _i, _p]
_L.b2s_toi_events.restype = C.c_long
_L.b2s_uniform53.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32]
_L.b2s_uniform53.restype = C.c_double

_f32 = np.float32


class b2Vec2:
    """pybox2d b2Vec2: float32 storage and float32 operator arithmetic."""
    __slots__ = ("_x", "_y")

    def __init__(self, x=0.0, y=None):
        if y is None:
            x, y = x
        self._x, self._y = _f32(x), _f32(y)

    x = property(lambda s: float(s._x))
    y = property(lambda s: float(s._y))

    def __iter__(self):
        yield float(self._x)
        yield float(self._y)

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (float(self._x), float(self._y))[i]

    def __mul__(self, a):      # b2Vec2.__mul__(float32 a)
        a = _f32(a)
        return b2Vec2(self._x * a, self._y * a)

    __rmul__ = __mul__          # b2Vec2.__rmul__(float32 a)

    def __neg__(self):
        return b2Vec2(-self._x, -self._y)

    def __add__(self, o):
        o = _vec(o)
        return b2Vec2(self._x + o._x, self._y + o._y)

    def __sub__(self, o):
        o = _vec(o)
        return b2Vec2(self._x - o._x, self._y - o._y)

    def __eq__(self, o):
        try:
            ox, oy = o
        except TypeError:
            return NotImplemented
        return float(self._x) == ox and float(self._y) == oy

    def __hash__(self):
        return hash((float(self._x), float(self._y)))

    def __repr__(self):
        return "b2Vec2(%r,%r)" % (float(self._x), float(self._y))


def _vec(v):
    return v if isinstance(v, b2Vec2) else b2Vec2(v[0], v[1])


def dot(a, b):                  # b2Dot: float32
    a, b = _vec(a), _vec(b)
    return float(_f32(a._x * b._x) + _f32(a._y * b._y))


class polygonShape:
    def __init__(self, box=None, vertices=None):
        self.box = None if box is None else tuple(box)
        self.verts = None if vertices is None else [tuple(v) for v in vertices]
        self._fixture = None   # (world, fixture id) once attached

    @property
    def vertices(self):          # b2PolygonShape.vertices: list of (x, y) tuples of the stored (hull-ordered) vertices
        w, fid = self._fixture
        buf = (C.c_float * 32)()
        n = _L.b2s_fixture_vertices(w._h, fid, buf)
        return [(float(buf[2 * i]), float(buf[2 * i + 1])) for i in range(n)]


class circleShape:               # imported by the reference, never instantiated
    pass


class fixtureDef:
    def __init__(self, shape=None, density=0.0, friction=0.2, restitution=0.0, userData=None, **kw):
        self.shape, self.density, self.friction, self.restitution, self.userData = shape, density, friction, restitution, userData


class contactListener:
    def __init__(self):
        pass

    def BeginContact(self, contact):
        pass

    def EndContact(self, contact):
        pass


staticBody, dynamicBody = 0, 2


class _Fixture:
    def __init__(self, body, fid, shape, userData=None):
        self.body, self._fid, self.shape, self.userData = body, fid, shape, userData
        shape._fixture = (body._w, fid)


class _Contact:
    def __init__(self, fa, fb):
        self.fixtureA, self.fixtureB = fa, fb


class b2Body:
    """A body handle.  Arbitrary attributes may be attached (the reference stores goal_contact on it, mrp00:377)."""

    def __init__(self, world, bid, userData):
        self._w, self._b, self.userData = world, bid, userData
        self._fixtures = []     # creation order

    def _get(self):
        buf = (C.c_float * 12)()
        _L.b2s_body_get(self._w._h, self._b, buf)
        return buf

    position = property(lambda s: b2Vec2(s._get()[0], s._get()[1]))
    angle = property(lambda s: float(s._get()[2]))
    worldCenter = property(lambda s: b2Vec2(s._get()[3], s._get()[4]))
    mass = property(lambda s: float(s._get()[8]))
    inertia = property(lambda s: float(s._get()[9]))
    localCenter = property(lambda s: b2Vec2(s._get()[10], s._get()[11]))

    @property
    def linearVelocity(self):
        g = self._get()
        return b2Vec2(g[5], g[6])

    @linearVelocity.setter
    def linearVelocity(self, v):
        _L.b2s_set_linear_velocity(self._w._h, self._b, float(v[0]), float(v[1]))

    @property
    def angularVelocity(self):
        return float(self._get()[7])

    @angularVelocity.setter
    def angularVelocity(self, w):
        if not isinstance(w, float):   # SWIG's float typemap (the reference casts with float(turn), mrp00:420)
            raise TypeError("in method 'b2Body___SetAngularVelocity', argument 2 of type 'float32'")
        _L.b2s_set_angular_velocity(self._w._h, self._b, float(w))

    @property
    def fixtures(self):          # b2Body.fixtures walks m_fixtureList: newest first
        return list(reversed(self._fixtures))

    def _attach(self, shape, density, friction, restitution, userData=None):
        h = self._w._h
        if shape.box is not None:
            b = shape.box
            if len(b) == 2:
                fid = _L.b2s_create_box_fixture(h, self._b, float(b[0]), float(b[1]), 0, 0, 0, 0, float(density), float(friction), float(restitution))
            else:
                fid = _L.b2s_create_box_fixture(h, self._b, float(b[0]), float(b[1]), 1, float(b[2][0]), float(b[2][1]), float(b[3]),
                                                float(density), float(friction), float(restitution))
        else:
            flat = np.asarray(shape.verts, dtype=np.float32).ravel()
            fid = _L.b2s_create_poly_fixture(h, self._b, flat.ctypes.data, len(shape.verts), float(density), float(friction), float(restitution))
            if fid < 0:
                raise ValueError("polygon needs 3..8 vertices")
        fx = _Fixture(self, fid, shape, userData)
        self._fixtures.append(fx)
        self._w._fix[fid] = fx
        return fx

    def CreatePolygonFixture(self, box=None, vertices=None, density=0.0, friction=0.2, restitution=0.0, **kw):
        return self._attach(polygonShape(box=box, vertices=vertices), density, friction, restitution)

    def ApplyForce(self, force, point, wake):
        _L.b2s_apply_force(self._w._h, self._b, float(force[0]), float(force[1]), float(point[0]), float(point[1]))

    def ApplyTorque(self, torque, wake):
        _L.b2s_apply_torque(self._w._h, self._b, float(torque))

    def ApplyLinearImpulse(self, impulse, point, wake):
        _L.b2s_apply_linear_impulse(self._w._h, self._b, float(impulse[0]), float(impulse[1]), float(point[0]), float(point[1]))

    def ApplyAngularImpulse(self, impulse, wake):
        _L.b2s_apply_angular_impulse(self._w._h, self._b, float(impulse))

    def GetWorldPoint(self, localPoint):
        out = (C.c_float * 2)()
        _L.b2s_world_point(self._w._h, self._b, float(localPoint[0]), float(localPoint[1]), out)
        return b2Vec2(out[0], out[1])

    def GetWorldVector(self, localVector):
        out = (C.c_float * 2)()
        _L.b2s_world_vector(self._w._h, self._b, float(localVector[0]), float(localVector[1]), out)
        return b2Vec2(out[0], out[1])


class b2World:
    def __init__(self, gravity=(0, 0), doSleep=True):
        assert tuple(gravity) == (0, 0) and doSleep is False, "refshim supports the reference's world settings only"
        self._h = _L.b2s_world_new()
        self._bodies, self._fix, self._live = {}, {}, 0
        self.contactListener = None

    def __del__(self):
        try:
            _L.b2s_world_free(self._h)
        except Exception:
            pass

    def _create(self, dynamic, position, angle, linearDamping, angularDamping, userData, fixtures):
        bid = _L.b2s_create_body(self._h, dynamic, float(position[0]), float(position[1]), float(angle), float(linearDamping), float(angularDamping))
        body = b2Body(self, bid, userData)
        self._bodies[bid] = body
        self._live += 1
        if fixtures is not None:
            for fd in (fixtures if isinstance(fixtures, (list, tuple)) else [fixtures]):
                body._attach(fd.shape, fd.density, fd.friction, fd.restitution, fd.userData)
        return body

    def CreateDynamicBody(self, position=(0, 0), angle=0.0, linearDamping=0.0, angularDamping=0.0, userData=None, fixtures=None, **kw):
        return self._create(1, position, angle, linearDamping, angularDamping, userData, fixtures)

    def CreateStaticBody(self, position=(0, 0), angle=0.0, userData=None, fixtures=None, **kw):
        return self._create(0, position, angle, 0.0, 0.0, userData, fixtures)

    def DestroyBody(self, body):
        """The reference only ever destroys every body of the world (with the listener detached), mrp00:218-229."""
        assert self.contactListener is None, "refshim: DestroyBody with a listener attached is not modelled"
        del self._bodies[body._b]
        self._live -= 1
        if self._live == 0:
            _L.b2s_world_clear(self._h)
            self._bodies, self._fix = {}, {}

    def Step(self, dt, velocityIterations, positionIterations):
        assert len(self._bodies) == self._live, "refshim: partial DestroyBody is not modelled"
        cap = 256
        ev = (C.c_int * (5 * cap))()
        n = _L.b2s_step(self._h, float(dt), velocityIterations, positionIterations, ev, cap)
        assert n <= cap
        lst = self.contactListener
        if lst is not None:      # callbacks in firing order (they only set flags, so deferring them past Step is equivalent)
            for k in range(n):
                begin, fa, fb = ev[5 * k], ev[5 * k + 3], ev[5 * k + 4]
                c = _Contact(self._fix[fa], self._fix[fb])
                (lst.BeginContact if begin else lst.EndContact)(c)

    @property
    def toi_events(self):
        return int(_L.b2s_toi_events(self._h))

    @property
    def contacts(self):
        """list of dicts, contact-list order (pybox2d: world.contacts / contact.manifold.points[j].normalImpulse)"""
        buf = (C.c_float * 12)()
        n = _L.b2s_contact_get(self._h, 0, buf)
        out = []
        for k in range(n):
            _L.b2s_contact_get(self._h, k, buf)
            pc = int(buf[3])
            out.append(dict(fixtureA=int(buf[0]), fixtureB=int(buf[1]), touching=bool(buf[2]), pointCount=pc,
                            points=[dict(localPoint=(buf[4 + 4 * j], buf[5 + 4 * j]), normalImpulse=buf[6 + 4 * j], tangentImpulse=buf[7 + 4 * j])
                                    for j in range(pc)]))
        return out

    def load_state(self, bodies6, fat4, contacts14):
        """harness only (no pybox2d counterpart): see b2s_load_state in oracle/b2shim_capi.cpp"""
        b = np.ascontiguousarray(bodies6, dtype=np.float32)
        f = np.ascontiguousarray(fat4, dtype=np.float32)
        c = np.ascontiguousarray(contacts14, dtype=np.uint32).reshape(-1, 14)
        rc = _L.b2s_load_state(self._h, b.shape[0], b.ctypes.data, f.shape[0], f.ctypes.data, c.shape[0], c.ctypes.data)
        assert rc == 0, rc


def uniform53(seed, stream, env, epoch, d):
    return float(_L.b2s_uniform53(seed, stream, env, epoch, d))
