// oracle/b2shim_capi.cpp — TEST INFRASTRUCTURE ONLY.
//
// A generic C view of the oracle's Box2D restatement (b2o::World in b2core.hpp): bodies, polygon fixtures,
// force / impulse / velocity accessors, Step with the Begin/EndContact events it fired.  tests/refshim/Box2D wraps
// it as a stand-in for the `Box2D` (pybox2d) module so that the REFERENCE'S OWN PYTHON env code
// (reference gym_puzzles/envs/multi_robot_puzzle_00.py / _02.py, imported unmodified from /root/reference) can run
// in this container.  That pins the oracle's restatement of the env logic (control law, distances, observation,
// reward, termination, reset order — oracle/mrp_env.hpp) against the reference itself; the Box2D arithmetic
// underneath stays a restatement (pybox2d is not installable here) and therefore unpinned.
//
// Entry points mirror the pybox2d calls the reference makes (SURVEY.md §8c call-site list).
#include <cstdint>
#include <cstring>
#include <vector>

#include "b2core.hpp"
#include "philox.hpp"

using namespace b2o;

namespace {
struct Event { int begin, bodyA, bodyB, fixA, fixB; };
struct ShimWorld : ContactListener {
    World* w = new World();
    std::vector<Event> events;
    ShimWorld() { w->listener = this; }
    ~ShimWorld() override { delete w; }
    void push(Contact* c, int begin) { events.push_back({begin, c->bA, c->bB, c->fA, c->fB}); }
    void BeginContact(Contact* c) override { push(c, 1); }
    void EndContact(Contact* c) override { push(c, 0); }
};
ShimWorld* S(void* h) { return (ShimWorld*)h; }
Body& B(void* h, int b) { return S(h)->w->bodies[b]; }
}  // namespace

extern "C" {

// Box2D.b2World(gravity=(0,0), doSleep=False)   reference mrp00:164, mrp02:149
void* b2s_world_new() { return new ShimWorld(); }
void b2s_world_free(void* h) { delete S(h); }
// every body destroyed (reference _destroy, mrp00:218-229): the world is empty again.  A fresh b2o::World is
// substituted (canonical proxy ids; the reference's reused world permutes them) and m_inv_dt0 is
// carried over as the real world would keep it.
void b2s_world_clear(void* h) {
    ShimWorld* s = S(h);
    const float inv_dt0 = s->w->inv_dt0;
    delete s->w;
    s->w = new World();
    s->w->listener = s;
    s->w->inv_dt0 = inv_dt0;
    s->events.clear();
}
// CreateDynamicBody / CreateStaticBody   mrp00:313-319,368-376,268-274; mrp02:322-328,363-389,402-410
int b2s_create_body(void* h, int dynamic, float x, float y, float angle, float lin_damp, float ang_damp) {
    return S(h)->w->CreateBody(dynamic ? kDynamic : kStatic, Vec2(x, y), angle, lin_damp, ang_damp);
}
// polygonShape(box=(hx, hy)) / polygonShape(box=(hx, hy, (cx, cy), angle))   mrp00:323-351,266; mrp02:332-341,375-383
int b2s_create_box_fixture(void* h, int body, float hx, float hy, int has_center, float cx, float cy, float angle,
                           float density, float friction, float restitution) {
    Polygon p;
    if (has_center) p.SetAsBox(hx, hy, Vec2(cx, cy), angle); else p.SetAsBox(hx, hy);
    return S(h)->w->CreateFixture(body, p, density, friction, restitution);
}
// polygonShape(vertices=[...])   mrp00:371, mrp02:367
int b2s_create_poly_fixture(void* h, int body, const float* xy, int n, float density, float friction, float restitution) {
    if (n < 3 || n > kMaxVerts) return -1;
    Vec2 pts[kMaxVerts];
    for (int i = 0; i < n; ++i) pts[i] = Vec2(xy[2 * i], xy[2 * i + 1]);
    Polygon p;
    p.Set(pts, n);
    return S(h)->w->CreateFixture(body, p, density, friction, restitution);
}
// fixture.shape.vertices   mrp00:356-361
int b2s_fixture_vertices(void* h, int fixture, float* xy) {
    const Polygon& p = S(h)->w->fixtures[fixture].shape;
    for (int i = 0; i < p.count; ++i) { xy[2 * i] = p.v[i].x; xy[2 * i + 1] = p.v[i].y; }
    return p.count;
}
int b2s_fixture_body(void* h, int fixture) { return S(h)->w->fixtures[fixture].body; }
// body.fixtures (newest first, as b2Body::m_fixtureList)
int b2s_body_fixtures(void* h, int body, int* out, int cap) {
    const Body& b = B(h, body);
    int n = 0;
    for (int f : b.fixtures) if (n < cap) out[n++] = f;
    return (int)b.fixtures.size();
}
// position, angle, worldCenter, linearVelocity, angularVelocity, mass, inertia, localCenter
void b2s_body_get(void* h, int body, float* out12) {
    const Body& b = B(h, body);
    out12[0] = b.xf.p.x; out12[1] = b.xf.p.y; out12[2] = b.sweep.a;
    out12[3] = b.sweep.c.x; out12[4] = b.sweep.c.y;
    out12[5] = b.v.x; out12[6] = b.v.y; out12[7] = b.w;
    out12[8] = b.mass; out12[9] = b.GetInertia();
    out12[10] = b.sweep.localCenter.x; out12[11] = b.sweep.localCenter.y;
}
// body.linearVelocity = ..., body.angularVelocity = ...   mrp00:419-420 (dynamic bodies only, as b2Body::Set*Velocity)
void b2s_set_linear_velocity(void* h, int body, float vx, float vy) { if (B(h, body).type == kDynamic) B(h, body).v = Vec2(vx, vy); }
void b2s_set_angular_velocity(void* h, int body, float w) { if (B(h, body).type == kDynamic) B(h, body).w = w; }
// ApplyForce / ApplyTorque / ApplyLinearImpulse / ApplyAngularImpulse   mrp00:424; mrp02:454,122,456,463-467,474
void b2s_apply_force(void* h, int body, float fx, float fy, float px, float py) { B(h, body).ApplyForce(Vec2(fx, fy), Vec2(px, py)); }
void b2s_apply_torque(void* h, int body, float t) { B(h, body).ApplyTorque(t); }
void b2s_apply_linear_impulse(void* h, int body, float ix, float iy, float px, float py) { B(h, body).ApplyLinearImpulse(Vec2(ix, iy), Vec2(px, py)); }
void b2s_apply_angular_impulse(void* h, int body, float imp) { B(h, body).ApplyAngularImpulse(imp); }
// GetWorldPoint / GetWorldVector   mrp00:471; mrp02:117,449-450
void b2s_world_point(void* h, int body, float x, float y, float* out2) { Vec2 p = B(h, body).GetWorldPoint(Vec2(x, y)); out2[0] = p.x; out2[1] = p.y; }
void b2s_world_vector(void* h, int body, float x, float y, float* out2) { Vec2 p = B(h, body).GetWorldVector(Vec2(x, y)); out2[0] = p.x; out2[1] = p.y; }
// world.Step(1.0/FPS, 6*30, 2*30)   mrp00:428, mrp02:478.  Returns the number of Begin/EndContact callbacks the step
// fired; events5 receives up to cap records {begin, bodyA, bodyB, fixtureA, fixtureB} in firing order.
int b2s_step(void* h, float dt, int vel_iters, int pos_iters, int* events5, int cap) {
    ShimWorld* s = S(h);
    s->events.clear();
    s->w->Step(dt, vel_iters, pos_iters);
    int n = 0;
    for (const Event& e : s->events) {
        if (n >= cap) break;
        int* o = events5 + 5 * n++;
        o[0] = e.begin; o[1] = e.bodyA; o[2] = e.bodyB; o[3] = e.fixA; o[4] = e.fixB;
    }
    return (int)s->events.size();
}
// Harness entry (no pybox2d counterpart): load a between-steps state — dynamic bodies (c, a, v, w), fat AABBs of their
// proxies, contact list in include/mrp_state.h's record format — so that the reference's Python can be stepped from
// the very state the oracle and the CUDA library are stepped from (BASELINE.json north_star: "from identical states").
// The env modules create the dynamic bodies (block, robots) before the walls, so canonical body / fixture ids are the
// world's creation-order ids.
int b2s_load_state(void* h, int n_dyn, const float* bodies6, int n_dynfix, const float* fat4, int nc, const uint32_t* contacts14) {
    World* w = S(h)->w;
    if (n_dyn > (int)w->bodies.size() || n_dynfix > (int)w->fixtures.size()) return -1;
    for (int b = 0; b < n_dyn; ++b) if (w->bodies[b].type != kDynamic) return -2;
    w->LoadState(n_dyn, bodies6, n_dynfix, fat4, nc, contacts14);
    return 0;
}
// contact k of the world's contact list (head first): {fixtureA, fixtureB, touching, pointCount, then per point
// localPoint.x, localPoint.y, normalImpulse, tangentImpulse}; returns the number of contacts.  For the known-answer tests
// (pybox2d: world.contacts[k].manifold.points[j].normalImpulse).
int b2s_contact_get(void* h, int k, float* out12) {
    World* w = S(h)->w;
    if (k >= 0 && k < (int)w->contactList.size()) {
        const Contact* c = w->contactList[k];
        out12[0] = (float)c->fA; out12[1] = (float)c->fB; out12[2] = c->touching ? 1.0f : 0.0f; out12[3] = (float)c->manifold.pointCount;
        for (int j = 0; j < 2; ++j) {
            const ManifoldPoint& p = c->manifold.points[j];
            out12[4 + 4 * j] = p.localPoint.x; out12[5 + 4 * j] = p.localPoint.y; out12[6 + 4 * j] = p.normalImpulse; out12[7 + 4 * j] = p.tangentImpulse;
        }
    }
    return (int)w->contactList.size();
}
long b2s_toi_events(void* h) { return S(h)->w->stat_toi_events; }
// Box2D version fork of the collision routine for the worlds of this library (b2core.hpp g_fork_230; no pybox2d counterpart:
// it stands for which box2d-py wheel is installed).  Returns the previous setting.
int b2s_set_box2d_fork(int fork_230) { int old = g_fork_230; g_fork_230 = fork_230 ? 1 : 0; return old; }

// the Philox stream the oracle / product spawn from, so the harness can feed the reference's np.random.uniform and
// action_space.sample() the very same draws (oracle/philox.hpp)
double b2s_uniform53(uint64_t seed, uint32_t stream, uint64_t env, uint32_t epoch, uint32_t d) {
    return orc::uniform53(seed, stream, env, epoch, d);
}

}  // extern "C"
