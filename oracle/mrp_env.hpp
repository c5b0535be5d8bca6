// oracle/mrp_env.hpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// CPU restatement of the Python env logic of gym_puzzles' MultiRobotPuzzle family:
//   reference gym_puzzles/envs/multi_robot_puzzle_00.py  (mrp00)  v0 / Heavy-v0
//   reference gym_puzzles/envs/multi_robot_puzzle_02.py  (mrp02)  v2 / Heavy-v2
// on top of oracle/b2core.hpp.  Python float arithmetic is float64, pybox2d b2Vec2
// arithmetic is float32; each expression below keeps the width the reference has.
// Pinning: this env logic IS checked against the reference's own Python, executed unmodified in the build container
// over a pybox2d stand-in backed by b2core.hpp (tests/refshim, tests/test_reference_python.py) and through the committed
// fixtures it produced (tests/golden/, tests/test_golden.py): flags and body state identical, observation / reward
// equal up to the float64 ulps of `**0.5` vs sqrt.  The Box2D arithmetic underneath is *** PARITY UNPINNED ***
// (see b2core.hpp header).
//
// Deliberate differences (DESIGN.md "Deviations"):
//  * spawns use counter-based Philox instead of numpy's global RNG (north star);
//  * a fresh b2World is built at every reset (proxy ids canonicalised);
//  * v2 decay^(-t) defaults to 1 when update_params was never called;
//  * TimeLimit (gym_puzzles/__init__.py:3-29) and gym-0.21 vector auto-reset are folded in.
#pragma once
#include <cmath>
#include <vector>

#include "../include/mrp_state.h"
#include "b2core.hpp"
#include "philox.hpp"

namespace orc {
using namespace b2o;

struct RewardParams {  // mrp00:231-239 / mrp02:216-225
    double agentDelta, agentDistance, blockDelta, blockDistance;
    double puzzleComp = 10000, outOfBounds = 1000, blkOutOfBounds = 100;
    double scaled_epsilon = 0.1;       // mrp02:161 (v2 only)
    double decay_pow = 1.0;            // decay**(-timestep), mrp02:227-230
};

inline double py_mod(double a, double b) {  // Python float %
    double r = std::fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}
inline double py_distance(double ax, double ay, double bx, double by) {  // mrp00:130-132
    // Python evaluates (x+y)**0.5 with libm pow(); IEEE sqrt is used here so that the value is
    // reproducible on every platform (the two differ by <= 1 ulp of float64 in rare cases).
    double x = (ax - bx) * (ax - bx), y = (ay - by) * (ay - by);
    return std::sqrt(x + y);
}

struct Env : ContactListener {
    int variant;
    bool v2, heavy;
    int n;  // robots
    mrp_layout L;
    uint64_t seed, gid;
    RewardParams rp;
    bool auto_reset = true;

    // module constants
    double SCALE, VIEW_W, VIEW_H, W, H;

    World* world = nullptr;
    // "square" extension (BASELINE.json configs[4], not a reference env): three blocks T, L, I with Heavy-v2 dynamics;
    // blocks[k] are world bodies 0..2, the goal block is blocks[min(blks_in_place, 2)] (block_queue order T, L, I,
    // mrp00:293-297) and block_body always names the current goal block.  With one block everything below reduces to the
    // reference's v2 env.
    bool square = false;
    int nblk = 1;
    int blocks[MRP_SQUARE_BLOCKS] = {0, 0, 0};
    std::vector<Vec2> sq_vertices[MRP_SQUARE_BLOCKS];  // per block, de-duplicated the reference's way (mrp00:356-361)
    double sq_target[MRP_SQUARE_BLOCKS][3];            // target COM offset from the goal centre (metres) and target angle
    int block_body = 0;
    std::vector<int> agent_body;
    bool goal_contact[MRP_MAX_AGENTS];
    Vec2 blk_vertices[8];  // mrp00:356-361 — bar verts then stem verts
    double agent_dist[MRP_MAX_AGENTS], block_distance;
    double goal_x, goal_y;
    int blks_in_place = 0, prev_blks_in_place = 0;
    int elapsed = 0, episode = -1;
    double ep_return = 0;
    int ep_len = 0;
    // stats
    long n_episodes = 0, n_success = 0, n_trunc = 0;
    double sum_return = 0, sum_len = 0;

    Env(int variant_, int n_agents, uint64_t seed_, uint64_t gid_) : variant(variant_), seed(seed_), gid(gid_) {
        mrp_layout_for(variant, n_agents, &L);
        n = L.n_agents;
        v2 = variant >= 2;
        heavy = (variant & 1) != 0;
        square = variant == MRP_VARIANT_SQUARE_V2;
        if (square) {
            heavy = true;                 // Heavy-v2 dynamics
            nblk = MRP_SQUARE_BLOCKS;
            // target poses of mrp00:83-88 (written there for block unit 0.5 m), in units of u = 0.1 m (the v2 T-block's unit):
            // COM of T at (0, 1.5u), of L at (-4/3 u, -4/3 u) turned by pi/2, of I at (2u, -u): together the square [-3u, 3u]^2
            const double u = 0.1;
            const double rel[3][3] = {{0.0, 0.75 / 0.5, 0.0}, {-2. / 3. / 0.5, -2. / 3. / 0.5, 0.5 * M_PI}, {1.0 / 0.5, -0.5 / 0.5, 0.0}};
            for (int k = 0; k < 3; ++k) { sq_target[k][0] = rel[k][0] * u; sq_target[k][1] = rel[k][1] * u; sq_target[k][2] = rel[k][2]; }
        }
        if (!v2) {
            SCALE = 30.0; VIEW_W = 640; VIEW_H = 480;                       // mrp00:40-42
            rp.agentDelta = 10; rp.agentDistance = 0.1; rp.blockDelta = 50; rp.blockDistance = 0.025;
            goal_x = 320.0 + 0.0 * SCALE;                                   // mrp00:115-128
            goal_y = 240.0 + 0.75 * SCALE;
        } else {
            SCALE = 140.0 * 4; VIEW_W = 1440; VIEW_H = 810;                 // mrp02:40-43
            rp.agentDelta = 10; rp.agentDistance = 0.25; rp.blockDelta = 25; rp.blockDistance = 0.1;
            goal_x = goal_y = 0;
        }
        W = VIEW_W / SCALE;
        H = VIEW_H / SCALE;
        for (int i = 0; i < MRP_MAX_AGENTS; ++i) { goal_contact[i] = false; agent_dist[i] = 0; }
        block_distance = 0;
        build_world();
    }
    ~Env() override { delete world; }
    Env(const Env&) = delete;

    // ---- ContactDetector (mrp00:92-111, mrp02:85-102)
    void contact_event(Contact* c, bool begin) {
        for (int i = 0; i < n; ++i) {
            int ab = agent_body[i];
            if (c->bA == ab || c->bB == ab)
                if (c->bA == block_body || c->bB == block_body) goal_contact[i] = begin;
        }
    }
    void BeginContact(Contact* c) override { contact_event(c, true); }
    void EndContact(Contact* c) override { contact_event(c, false); }

    // ---- world construction with given poses (fresh world, canonical proxy ids)
    void build_world(const double* blk = nullptr, const double* agents = nullptr) {
        delete world;
        world = new World();
        world->listener = this;
        agent_body.clear();
        float damp = 5.0f;
        if (square) {
            // blocks.py:70-109 / mrp00:320-351 shapes at unit u = 0.1, Heavy-v2 material (mrp02:320-342); blk: [3][3] poses
            const float u = 0.1f;
            for (int k = 0; k < 3; ++k) {
                double bx = blk ? blk[3 * k] : W / 2, by = blk ? blk[3 * k + 1] : H * (k + 1) / 4, ba = blk ? blk[3 * k + 2] : 0.0;
                blocks[k] = world->CreateBody(kDynamic, Vec2((float)bx, (float)by), (float)ba, damp, damp);
                Polygon a, b;
                std::vector<Polygon> fx;   // creation order
                if (k == 0) { a.SetAsBox(1 * u, 1 * u, Vec2(0.0f, -1 * u), 0.0f); b.SetAsBox(3 * u, 1 * u, Vec2(0.0f, 1 * u), 0.0f); fx = {a, b}; }
                else if (k == 1) { a.SetAsBox(1 * u, 1 * u, Vec2(1 * u, 0.5f * u), 0.0f); b.SetAsBox(1 * u, 2 * u, Vec2(-1 * u, -0.5f * u), 0.0f); fx = {a, b}; }
                else { a.SetAsBox(1 * u, 2 * u); fx = {a}; }
                for (const Polygon& p : fx) world->CreateFixture(blocks[k], p, 20.0f, 0.01f, 0.0f);
                // vertex list: fixtures newest first, vertices not seen before (mrp00:356-361)
                sq_vertices[k].clear();
                for (int f = (int)fx.size() - 1; f >= 0; --f)
                    for (int i = 0; i < fx[f].count; ++i) {
                        bool seen = false;
                        for (const Vec2& q : sq_vertices[k]) if (q.x == fx[f].v[i].x && q.y == fx[f].v[i].y) seen = true;
                        if (!seen) sq_vertices[k].push_back(fx[f].v[i]);
                    }
            }
            block_body = blocks[blks_in_place < 2 ? (blks_in_place < 0 ? 0 : blks_in_place) : 2];
        } else {
        double bx = blk ? blk[0] : W / 2, by = blk ? blk[1] : H / 2, ba = blk ? blk[2] : 0.0;
        // _generate_blocks  (mrp00:299-361, mrp02:313-350)
        block_body = world->CreateBody(kDynamic, Vec2((float)bx, (float)by), (float)ba, damp, damp);
        blocks[0] = block_body;
        Polygon stem, bar;
        if (!v2) {
            double S = 2.0;
            double scaled = heavy ? S / 2 : S;
            double dense = heavy ? 5.0 * 2 : 5.0;
            stem.SetAsBox((float)(1 / scaled), (float)(1 / scaled), Vec2(0.0f, (float)(-1 / scaled)), 0.0f);
            bar.SetAsBox((float)(3 / scaled), (float)(1 / scaled), Vec2(0.0f, (float)(1 / scaled)), 0.0f);
            world->CreateFixture(block_body, stem, (float)dense, 0.999f, 0.0f);
            world->CreateFixture(block_body, bar, (float)dense, 0.999f, 0.0f);
        } else {
            double dense = heavy ? 20.0 : 1.56;
            stem.SetAsBox(0.1f, 0.1f, Vec2(0.0f, -0.1f), 0.0f);
            bar.SetAsBox(0.3f, 0.1f, Vec2(0.0f, 0.1f), 0.0f);
            world->CreateFixture(block_body, stem, (float)dense, 0.01f, 0.0f);
            world->CreateFixture(block_body, bar, (float)dense, 0.01f, 0.0f);
        }
        // blks_vertices: block.fixtures iterates newest-first => bar, then stem
        for (int i = 0; i < 4; ++i) { blk_vertices[i] = bar.v[i]; blk_vertices[4 + i] = stem.v[i]; }
        }
        // _generate_agents  (mrp00:363-378, mrp02:352-392)
        for (int i = 0; i < n; ++i) {
            double ax = agents ? agents[2 * i] : 1.0 + i, ay = agents ? agents[2 * i + 1] : 1.0;
            Vec2 pts[8];
            Polygon oct;
            if (!v2) {
                const double S = 2.0;
                const double P[8][2] = {{-0.5 / S, -1.5 / S}, {0.5 / S, -1.5 / S}, {1.5 / S, -0.5 / S}, {1.5 / S, 0.5 / S},
                                        {0.5 / S, 1.5 / S},   {-0.5 / S, 1.5 / S}, {-1.5 / S, 0.5 / S}, {-1.5 / S, -0.5 / S}};
                for (int k = 0; k < 8; ++k) pts[k] = Vec2((float)P[k][0], (float)P[k][1]);
                oct.Set(pts, 8);
                int b = world->CreateBody(kDynamic, Vec2((float)ax, (float)ay), 0.0f, damp, damp);
                world->CreateFixture(b, oct, 0.0f, 0.2f, 0.0f);  // fixtureDef defaults
                agent_body.push_back(b);
            } else {
                const double P[8][2] = {{-0.039, -0.095}, {0.039, -0.095}, {0.095, -0.039}, {0.095, 0.039},
                                        {0.039, 0.095},   {-0.039, 0.095}, {-0.095, 0.039}, {-0.095, -0.039}};
                for (int k = 0; k < 8; ++k) pts[k] = Vec2((float)P[k][0], (float)P[k][1]);
                oct.Set(pts, 8);
                double theta = 3.0 / 2 * M_PI;  // SIMPLE, mrp02:360
                int b = world->CreateBody(kDynamic, Vec2((float)ax, (float)ay), (float)theta, damp, damp);
                world->CreateFixture(b, oct, 17.3f, 0.01f, 0.0f);
                Polygon w1, w2;
                w1.SetAsBox(0.005f, 0.05f, Vec2(0.06f, 0.0f), 0.0f);
                w2.SetAsBox(0.005f, 0.05f, Vec2(-0.06f, 0.0f), 0.0f);
                world->CreateFixture(b, w1, 0.0f, 0.01f, 0.0f);
                world->CreateFixture(b, w2, 0.0f, 0.01f, 0.0f);
                agent_body.push_back(b);
            }
        }
        // _generate_boundary (mrp00:260-275, mrp02:394-411)
        const double borders[4][2] = {{0, 0.5}, {1, 0.5}, {0.5, 0}, {0.5, 1}};
        for (int i = 0; i < 4; ++i) {
            double hx, hy;
            if (!v2) { if (i < 2) { hx = 1.0; hy = H; } else { hx = W; hy = 1.0; } }
            else     { if (i < 2) { hx = 0.1; hy = H; } else { hx = W; hy = 0.1; } }
            int b = world->CreateBody(kStatic, Vec2((float)(W * borders[i][0]), (float)(H * borders[i][1])), 0.0f, 0.0f, 0.0f);
            Polygon box;
            box.SetAsBox((float)hx, (float)hy);
            world->CreateFixture(b, box, 0.0f, 0.2f, 0.0f);
        }
        for (int i = 0; i < MRP_MAX_AGENTS; ++i) goal_contact[i] = false;
    }

    Vec2 wc(int body) const { return world->bodies[body].sweep.c; }  // worldCenter

    void calculate_distance() {  // mrp00:277-283 / mrp02:263-269
        Vec2 c = wc(block_body);
        if (!v2) {
            float px = c.x * (float)SCALE, py = c.y * (float)SCALE;  // b2Vec2 * float: float32
            block_distance = py_distance((double)px, (double)py, goal_x, goal_y);
        } else {
            double ratio = SCALE / VIEW_W;
            block_distance = py_distance((double)c.x * ratio, (double)c.y * ratio, target_x(goal_index()), target_y(goal_index()));
        }
    }
    // square: index of the current goal block and the target COM of block k (goal centre + offset, in obs units)
    int goal_index() const { return square ? (blks_in_place < 2 ? blks_in_place : 2) : 0; }
    double target_x(int k) const { return square ? goal_x + sq_target[k][0] * (SCALE / VIEW_W) : goal_x; }
    double target_y(int k) const { return square ? goal_y + sq_target[k][1] * (SCALE / VIEW_W) : goal_y; }
    void calculate_agent_distance() {  // mrp00:285-291 / mrp02:271-277
        Vec2 b = wc(block_body);
        for (int i = 0; i < n; ++i) {
            Vec2 a = wc(agent_body[i]);
            if (!v2) {
                float s = (float)SCALE;
                agent_dist[i] = py_distance((double)(a.x * s), (double)(a.y * s), (double)(b.x * s), (double)(b.y * s));
            } else {
                double ratio = SCALE / VIEW_W;
                agent_dist[i] = py_distance((double)a.x * ratio, (double)a.y * ratio, (double)b.x * ratio, (double)b.y * ratio);
            }
        }
    }

    // ---- spawn (mrp00:392-409 / mrp02:421-440) with Philox instead of np.random
    void spawn() {
        uint32_t d = 0;
        auto U = [&](double lo, double hi) { return lo + (hi - lo) * uniform53(seed, kStreamSpawn, gid, (uint32_t)episode, d++); };
        double blk[3 * MRP_SQUARE_BLOCKS], ag[2 * MRP_MAX_AGENTS];
        if (square) {
            // extension: the three blocks start on the vertical centre line, each with its own random angle; robots and goal as v2
            const double BORDER = 0.3;
            blks_in_place = 0;   // the block queue starts over (mrp00:403-404)
            for (int k = 0; k < 3; ++k) { blk[3 * k] = W / 2; blk[3 * k + 1] = H * (k + 1) / 4; blk[3 * k + 2] = U(0, 2 * M_PI); }
            for (int i = 0; i < n; ++i) { ag[2 * i] = U(BORDER, W / 3 - BORDER); ag[2 * i + 1] = U(BORDER, H - BORDER); }
        } else if (!v2) {
            const double BORDER = 1;
            blk[0] = U(BORDER, W - BORDER);
            blk[1] = U(BORDER, H - BORDER);
            blk[2] = U(0, 2 * M_PI);
            for (int i = 0; i < n; ++i) { ag[2 * i] = U(BORDER, W - BORDER); ag[2 * i + 1] = U(BORDER, H - BORDER); }
        } else {
            const double BORDER = 0.3;
            blk[0] = W / 2; blk[1] = H / 2;
            blk[2] = U(0, 2 * M_PI);
            for (int i = 0; i < n; ++i) { ag[2 * i] = U(BORDER, W / 3 - BORDER); ag[2 * i + 1] = U(BORDER, H - BORDER); }
        }
        build_world(blk, ag);
        if (v2) {  // _set_random_goal mrp02:303-311
            const double B = 0.4;
            double x = U(W * 2 / 3 + B, W - B);
            double y = U(B, H - B);
            double ratio = SCALE / VIEW_W;
            goal_x = x * ratio;
            goal_y = y * ratio;
        }
        calculate_distance();
        calculate_agent_distance();
    }
    void hidden_action(float* a) const {  // action_space.sample() at mrp00:411 / mrp02:442
        for (int i = 0; i < L.act_dim; ++i)
            a[i] = (float)(-1.0 + 2.0 * uniform53(seed, kStreamResetAction, gid, (uint32_t)episode, (uint32_t)i));
    }

    // ---- env.step body (mrp00:413-521 / mrp02:444-584); obs: double[obs_dim]
    void env_step(const float* action, double* obs, double* reward_out, bool* done_out) {
        Body& blk = world->bodies[block_body];
        if (!v2) {
            const double SPEED = 10 / SCALE * 4;  // mrp00:50
            for (int i = 0; i < n; ++i) {
                Body& ag = world->bodies[agent_body[i]];
                float x = action[3 * i], y = action[3 * i + 1], turn = action[3 * i + 2];
                ag.v = Vec2((float)((double)x * SPEED), (float)((double)y * SPEED));
                ag.w = turn;
                double force = std::pow(1.1, -agent_dist[i]);
                double Ax = ag.sweep.c.x, Ay = ag.sweep.c.y, Bx = blk.sweep.c.x, By = blk.sweep.c.y;
                double denom = std::max(std::fabs(Bx - Ax), std::fabs(By - Ay));
                double sx = (Bx - Ax) / denom, sy = (By - Ay) / denom;
                blk.ApplyForce(Vec2((float)(force * sx), (float)(force * sy)), blk.sweep.c);
            }
        } else {
            const double FORCE = 0.75;
            for (int i = 0; i < n; ++i) {
                Body& ag = world->bodies[agent_body[i]];
                float turn = action[2 * i], vel = action[2 * i + 1];
                Vec2 f = ag.GetWorldVector(Vec2(0.0f, 1.0f));
                Vec2 p = ag.GetWorldPoint(Vec2(0.0f, 2.0f));
                Vec2 ff((float)((double)f.x * (double)vel * FORCE), (float)((double)f.y * (double)vel * FORCE));
                ag.ApplyForce(ff, p);
                // updateFriction (mrp02:116-122): all b2Vec2 float32 arithmetic
                Vec2 rn = ag.GetWorldVector(Vec2(1.0f, 0.0f));
                float dn = Dot(rn, ag.v);
                Vec2 lat = dn * rn;
                Vec2 imp = ag.mass * (-lat);
                ag.ApplyLinearImpulse(imp, ag.sweep.c);
                ag.ApplyAngularImpulse((float)(0.1 * (double)ag.GetInertia() * (double)ag.w));
                double max_torque = 0.0005;
                double torque = (double)std::fabs(turn) * max_torque;
                float tturn = turn;
                if (std::fabs(vel) < 0.1) tturn = 0;   // np.float32 abs vs python 0.1 (float64 compare)
                if (tturn < 0) ag.ApplyTorque((float)torque);
                else if (tturn > 0) ag.ApplyTorque((float)-torque);
                else ag.ApplyTorque(0.0f);
                double force = std::pow(10.0, -agent_dist[i]);
                force /= 50;
                double Ax = ag.sweep.c.x, Ay = ag.sweep.c.y, Bx = blk.sweep.c.x, By = blk.sweep.c.y;
                double denom = std::max(std::fabs(Bx - Ax), std::fabs(By - Ay));
                double sx = (Bx - Ax) / denom, sy = (By - Ay) / denom;
                blk.ApplyForce(Vec2((float)(force * sx), (float)(force * sy)), blk.sweep.c);
            }
        }
        world->Step((float)(1.0 / 50), 6 * 30, 2 * 30);

        double prev_agent_dist[MRP_MAX_AGENTS];
        for (int i = 0; i < n; ++i) prev_agent_dist[i] = agent_dist[i];
        double prev_distance = block_distance;
        calculate_distance();
        calculate_agent_distance();

        int o = 0;
        bool in_place = false;
        double reward = 0;
        if (!v2) {
            for (int i = 0; i < n; ++i) {
                Vec2 a = wc(agent_body[i]), b = wc(block_body);
                obs[o++] = (double)a.x * SCALE - (double)b.x * SCALE;
                obs[o++] = (double)a.y * SCALE - (double)b.y * SCALE;
                obs[o++] = agent_dist[i];
                obs[o++] = goal_contact[i] ? 1.0 : 0.0;
            }
            Vec2 c = wc(block_body);
            double x = c.x, y = c.y;
            double angle = py_mod((double)blk.sweep.a, 2 * M_PI);
            double fx = goal_x, fy = goal_y;
            x *= SCALE; y *= SCALE;
            double a_diff = py_mod(0.0, 2 * M_PI) - angle;
            in_place = !(std::fabs(fx - x) > 25.0) && !(std::fabs(fy - y) > 25.0);
            obs[o++] = x - fx; obs[o++] = y - fy; obs[o++] = a_diff;
            obs[o++] = py_distance(x, y, fx, fy);
            for (int k = 0; k < 8; ++k) {
                Vec2 p = blk.GetWorldPoint(blk_vertices[k]);
                obs[o++] = (double)p.x * SCALE;
                obs[o++] = (double)p.y * SCALE;
            }
            const double DS = 1.0;
            double deltaDist = prev_distance - block_distance;
            reward += deltaDist * rp.blockDelta * DS / 4.;
            reward -= rp.blockDistance * block_distance * DS / 4.;
            for (int i = 0; i < n; ++i) {
                double deltaAgent = prev_agent_dist[i] - agent_dist[i];
                reward += deltaAgent * rp.agentDelta * DS / 4.;
                reward -= rp.agentDistance * agent_dist[i] * DS / 4.;
                if (goal_contact[i]) reward += 0.25;
            }
            bool done = false;
            prev_blks_in_place = blks_in_place;
            blks_in_place = in_place ? 1 : 0;
            reward += (blks_in_place - prev_blks_in_place) * 10;
            if (blks_in_place == 1) { done = true; reward += 10000; }
            *reward_out = reward;
            *done_out = done;
            return;
        }
        if (square) { square_post(obs, prev_agent_dist, prev_distance, reward_out, done_out); return; }
        // ---- v2 (mrp02:491-584)
        double ratio = SCALE / VIEW_W;
        for (int i = 0; i < n; ++i) {
            const Body& ag = world->bodies[agent_body[i]];
            double aX = (double)ag.sweep.c.x * ratio, aY = (double)ag.sweep.c.y * ratio;
            double theta = py_mod((double)ag.sweep.a, 2 * M_PI);
            double norm_theta = theta <= M_PI ? -theta / M_PI : (2 * M_PI - theta) / M_PI;
            obs[o++] = aX; obs[o++] = aY; obs[o++] = norm_theta;
            double bX = (double)blk.sweep.c.x * ratio, bY = (double)blk.sweep.c.y * ratio;
            obs[o++] = aX - bX; obs[o++] = aY - bY;
            obs[o++] = ag.v.x; obs[o++] = ag.v.y; obs[o++] = ag.w;
            obs[o++] = agent_dist[i];
        }
        {
            double x = (double)blk.sweep.c.x * ratio, y = (double)blk.sweep.c.y * ratio;
            double angle = py_mod((double)blk.sweep.a, 2 * M_PI);
            double fx = goal_x, fy = goal_y;
            double a_diff = py_mod(0.0, 2 * M_PI) - angle;
            a_diff /= M_PI;
            in_place = !(std::fabs(fx - x) > rp.scaled_epsilon) && !(std::fabs(fy - y) > rp.scaled_epsilon);
            obs[o++] = x - fx; obs[o++] = y - fy; obs[o++] = a_diff;
            obs[o++] = py_distance(x, y, fx, fy);
            for (int k = 0; k < 8; ++k) {
                Vec2 p = blk.GetWorldPoint(blk_vertices[k]);
                obs[o++] = (double)p.x * ratio;
                obs[o++] = (double)p.y * ratio;
            }
        }
        obs[o++] = rp.scaled_epsilon;
        double deltaDist = prev_distance - block_distance;
        reward += deltaDist * rp.blockDelta;
        reward -= rp.blockDistance * block_distance;
        for (int i = 0; i < n; ++i) {
            double deltaAgent = prev_agent_dist[i] - agent_dist[i];
            reward += deltaAgent * rp.agentDelta;
            reward -= rp.agentDistance * agent_dist[i];
        }
        const double BOUNDS = 0.1;
        auto oob = [&](Vec2 c) {
            double x = c.x, y = c.y;
            if (x < BOUNDS || x > (W - BOUNDS)) return true;
            if (y < BOUNDS || y > (H - BOUNDS)) return true;
            return false;
        };
        bool agt_oob = false;
        for (int i = 0; i < n; ++i) if (oob(wc(agent_body[i]))) { agt_oob = true; break; }
        if (agt_oob) {
            reward -= rp.outOfBounds * rp.decay_pow;
            *reward_out = reward; *done_out = true;
            return;
        }
        if (oob(wc(block_body))) {
            reward -= rp.blkOutOfBounds * rp.decay_pow;
            *reward_out = reward; *done_out = true;
            return;
        }
        prev_blks_in_place = blks_in_place;
        blks_in_place = in_place ? 1 : 0;
        int num_in_contact = 0;
        for (int i = 0; i < n; ++i) if (goal_contact[i]) ++num_in_contact;
        bool done = false;
        if (blks_in_place == 1) {
            done = true;
            reward += rp.puzzleComp * rp.decay_pow * ((double)num_in_contact / (double)n);
        }
        *reward_out = reward;
        *done_out = done;
    }

    // ---- square extension: observation / reward / termination after world.Step, modelled on mrp02:491-584.
    // obs: per robot the 9 values of v2 (relative to the goal block); per block k = T, L, I: (x - tx, y - ty, (ta - angle) / pi,
    // distance to its target, vertices); epsilon; goal-block index; blocks in place; robots touching the goal block.  Reward: v2's shaping on the goal block;
    // when the goal block's COM is within epsilon of its target: + puzzleComp * decay * contacts / n, the next block of the
    // queue becomes the goal (distances re-based on it, contact flags cleared); done when all three are placed.
    void square_post(double* obs, const double* prev_agent_dist, double prev_distance, double* reward_out, bool* done_out) {
        const double ratio = SCALE / VIEW_W;
        const int g = goal_index();
        const Body& blk = world->bodies[blocks[g]];
        int o = 0;
        double reward = 0;
        for (int i = 0; i < n; ++i) {
            const Body& ag = world->bodies[agent_body[i]];
            double aX = (double)ag.sweep.c.x * ratio, aY = (double)ag.sweep.c.y * ratio;
            double theta = py_mod((double)ag.sweep.a, 2 * M_PI);
            double norm_theta = theta <= M_PI ? -theta / M_PI : (2 * M_PI - theta) / M_PI;
            obs[o++] = aX; obs[o++] = aY; obs[o++] = norm_theta;
            double bX = (double)blk.sweep.c.x * ratio, bY = (double)blk.sweep.c.y * ratio;
            obs[o++] = aX - bX; obs[o++] = aY - bY;
            obs[o++] = ag.v.x; obs[o++] = ag.v.y; obs[o++] = ag.w;
            obs[o++] = agent_dist[i];
        }
        bool in_place = false;
        for (int k = 0; k < 3; ++k) {
            const Body& bk = world->bodies[blocks[k]];
            double x = (double)bk.sweep.c.x * ratio, y = (double)bk.sweep.c.y * ratio;
            double angle = py_mod((double)bk.sweep.a, 2 * M_PI);
            double fx = target_x(k), fy = target_y(k);
            double a_diff = (sq_target[k][2] - angle) / M_PI;
            if (k == g) in_place = !(std::fabs(fx - x) > rp.scaled_epsilon) && !(std::fabs(fy - y) > rp.scaled_epsilon);
            obs[o++] = x - fx; obs[o++] = y - fy; obs[o++] = a_diff;
            obs[o++] = py_distance(x, y, fx, fy);
            for (const Vec2& v : sq_vertices[k]) {
                Vec2 p = bk.GetWorldPoint(v);
                obs[o++] = (double)p.x * ratio;
                obs[o++] = (double)p.y * ratio;
            }
        }
        obs[o++] = rp.scaled_epsilon;
        obs[o++] = (double)g;
        obs[o++] = (double)blks_in_place;
        {
            int touching = 0;
            for (int i = 0; i < n; ++i) if (goal_contact[i]) ++touching;
            obs[o++] = (double)touching;
        }
        reward += (prev_distance - block_distance) * rp.blockDelta;
        reward -= rp.blockDistance * block_distance;
        for (int i = 0; i < n; ++i) {
            reward += (prev_agent_dist[i] - agent_dist[i]) * rp.agentDelta;
            reward -= rp.agentDistance * agent_dist[i];
        }
        const double BOUNDS = 0.1;
        auto oob = [&](Vec2 c) {
            double x = c.x, y = c.y;
            return (x < BOUNDS || x > (W - BOUNDS)) || (y < BOUNDS || y > (H - BOUNDS));
        };
        for (int i = 0; i < n; ++i)
            if (oob(wc(agent_body[i]))) { *reward_out = reward - rp.outOfBounds * rp.decay_pow; *done_out = true; return; }
        for (int k = 0; k < 3; ++k)
            if (oob(wc(blocks[k]))) { *reward_out = reward - rp.blkOutOfBounds * rp.decay_pow; *done_out = true; return; }
        bool done = false;
        if (in_place) {
            int num_in_contact = 0;
            for (int i = 0; i < n; ++i) if (goal_contact[i]) ++num_in_contact;
            reward += rp.puzzleComp * rp.decay_pow * ((double)num_in_contact / (double)n);
            prev_blks_in_place = blks_in_place;
            ++blks_in_place;
            if (blks_in_place == 3) done = true;
            else {  // _set_next_goal_block
                block_body = blocks[goal_index()];
                for (int i = 0; i < n; ++i) goal_contact[i] = false;
                calculate_distance();
                calculate_agent_distance();
            }
        }
        *reward_out = reward;
        *done_out = done;
    }

    // ---- reset(): respawn + hidden random-action step (mrp00:392-411)
    void reset(double* obs) {
        ++episode;
        elapsed = 0;
        ep_return = 0;
        ep_len = 0;
        spawn();
        float a[3 * MRP_MAX_AGENTS];
        hidden_action(a);
        double r;
        bool d;
        env_step(a, obs, &r, &d);
    }

    // ---- TimeLimit + vector auto-reset step
    void step(const float* action, double* obs, double* reward, uint8_t* done, uint8_t* trunc) {
        bool d;
        env_step(action, obs, reward, &d);
        ++elapsed;
        bool limit = elapsed >= L.max_episode_steps;
        *trunc = (limit && !d) ? 1 : 0;
        *done = (d || limit) ? 1 : 0;
        ep_return += *reward;
        ep_len += 1;
        if (*done) {
            ++n_episodes;
            if (d) ++n_success;  // for v2 "d" also covers OOB; split out below if needed
            if (*trunc) ++n_trunc;
            sum_return += ep_return;
            sum_len += ep_len;
            if (auto_reset) reset(obs);
        }
    }

    // ---- canonical state exchange (include/mrp_state.h)
    void get_state(uint32_t* w) const {
        std::memset(w, 0, sizeof(uint32_t) * L.state_words);
        w[0] = (uint32_t)elapsed;
        w[1] = (uint32_t)episode;
        w[2] = (uint32_t)blks_in_place;
        w[3] = (uint32_t)world->contactList.size();
        for (int i = 0; i < n; ++i) w[L.off_goal_contact + i] = goal_contact[i] ? 1 : 0;
        float* fb = (float*)(w + L.off_bodies);
        for (int b = 0; b < nblk + n; ++b) {
            const Body& B = world->bodies[b < nblk ? blocks[b] : agent_body[b - nblk]];
            fb[6 * b + 0] = B.sweep.c.x; fb[6 * b + 1] = B.sweep.c.y; fb[6 * b + 2] = B.sweep.a;
            fb[6 * b + 3] = B.v.x; fb[6 * b + 4] = B.v.y; fb[6 * b + 5] = B.w;
        }
        double dd[MRP_MAX_AGENTS + 1];
        for (int i = 0; i < n; ++i) dd[i] = agent_dist[i];
        dd[n] = block_distance;
        std::memcpy(w + L.off_dists, dd, sizeof(double) * (n + 1));
        double g[2] = {goal_x, goal_y};
        std::memcpy(w + L.off_goal, g, sizeof(g));
        std::memcpy(w + L.off_episode_acc, &ep_return, sizeof(double));
        w[L.off_episode_acc + 2] = (uint32_t)ep_len;
        float* fa = (float*)(w + L.off_aabb);
        for (int f = 0; f < L.n_dyn_fixtures; ++f) {
            fa[4 * f + 0] = world->fat[f].lo.x; fa[4 * f + 1] = world->fat[f].lo.y;
            fa[4 * f + 2] = world->fat[f].hi.x; fa[4 * f + 3] = world->fat[f].hi.y;
        }
        int nc = std::min((int)world->contactList.size(), L.max_contacts);
        for (int k = 0; k < nc; ++k) {
            const Contact* c = world->contactList[k];
            uint32_t* cw = w + L.off_contacts + MRP_CONTACT_WORDS * k;
            const Manifold& m = c->manifold;
            // words beyond pointCount hold stale data in Box2D (b2Manifold is not cleared); they are
            // never read again, so the canonical record zeroes them
            int pc = m.pointCount;
            uint32_t type = pc > 0 ? (uint32_t)m.type : 0u;
            cw[0] = (uint32_t)c->fA | ((uint32_t)c->fB << 8) | ((c->touching ? 1u : 0u) << 16) | (type << 17) | ((uint32_t)pc << 18);
            if (pc == 0) continue;
            auto key16 = [](const ContactID& id) { return (uint32_t)(id.indexA | (id.indexB << 4) | (id.typeA << 8) | (id.typeB << 9)); };
            cw[1] = key16(m.points[0].id) | (pc > 1 ? (key16(m.points[1].id) << 16) : 0u);
            float* cf = (float*)cw;
            cf[2] = m.localNormal.x; cf[3] = m.localNormal.y; cf[4] = m.localPoint.x; cf[5] = m.localPoint.y;
            for (int j = 0; j < pc; ++j) {
                cf[6 + 4 * j] = m.points[j].localPoint.x; cf[7 + 4 * j] = m.points[j].localPoint.y;
                cf[8 + 4 * j] = m.points[j].normalImpulse; cf[9 + 4 * j] = m.points[j].tangentImpulse;
            }
        }
    }
    void set_state(const uint32_t* w) {
        elapsed = (int)w[0];
        episode = (int)w[1];
        blks_in_place = (int)w[2];
        int nc = (int)w[3];
        const float* fb = (const float*)(w + L.off_bodies);
        double blk[3 * MRP_SQUARE_BLOCKS] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ag[2 * MRP_MAX_AGENTS];
        for (int i = 0; i < 2 * MRP_MAX_AGENTS; ++i) ag[i] = 0;
        build_world(blk, ag);  // shapes/masses; poses overwritten below (blks_in_place is set: the goal block is chosen by it)
        // canonical body b is world body b (block(s), then the robots; walls are created last)
        world->LoadState(nblk + n, fb, L.n_dyn_fixtures, (const float*)(w + L.off_aabb), nc, w + L.off_contacts);
        for (int i = 0; i < n; ++i) goal_contact[i] = w[L.off_goal_contact + i] != 0;
        double dd[MRP_MAX_AGENTS + 1];
        std::memcpy(dd, w + L.off_dists, sizeof(double) * (n + 1));
        for (int i = 0; i < n; ++i) agent_dist[i] = dd[i];
        block_distance = dd[n];
        double g[2];
        std::memcpy(g, w + L.off_goal, sizeof(g));
        goal_x = g[0]; goal_y = g[1];
        std::memcpy(&ep_return, w + L.off_episode_acc, sizeof(double));
        ep_len = (int)w[L.off_episode_acc + 2];
    }
};

}  // namespace orc
