// mrp_env.cuh — the env-level half of the hot path, per lane: control laws, observation,
// reward, termination, TimeLimit, Philox spawn.  Mirrors
//   reference gym_puzzles/envs/multi_robot_puzzle_00.py:413-521 (v0 / Heavy-v0 step),
//   reference gym_puzzles/envs/multi_robot_puzzle_02.py:444-584 (v2 / Heavy-v2 step),
//   reset + _generate_* (mrp00:299-411, mrp02:303-442).
// Python float arithmetic is float64 and pybox2d b2Vec2 arithmetic is float32; every
// expression keeps the width it has in the reference.
#pragma once
#include "mrp_sim.cuh"

namespace mrp {

MRP_HDN double py_mod(double a, double b) {  // Python float % (not inlined: float64 fmod is ~300 instructions)
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}
MRP_HDN double py_distance(double ax, double ay, double bx, double by) {  // mrp00:130-132 (not inlined: float64 sqrt)
    double x = (ax - bx) * (ax - bx), y = (ay - by) * (ay - by);
    return sqrt(x + y);
}
constexpr double kTwoPiD = 2.0 * 3.141592653589793;
constexpr double kPiD = 3.141592653589793;

struct Env : Sim {
    MRP_HD Env(const SimConst& k, float* sm_, const float* ct_, int64_t env, float* vc_ = nullptr, int fdyn_ = kDynFields)
        : Sim(k, sm_, ct_, env, vc_, fdyn_) {}

    // b2Island::Solve's velocity integration + damping (A.8) for body b with accumulated force/torque.
    // Nothing between the control block and the island solve reads velocities, so doing it here is
    // bit-identical to doing it inside Solve.
    MRP_HD void integrate_velocity(int b, V2 force, float torque) {
        V2 v = mk(B(b, 3), B(b, 4));
        float w = B(b, 5);
        float im = invMass(b), ii = invI(b);
        v = v + K.h * (im * force);
        w += K.h * ii * torque;
        v = K.lin_k * v;
        w *= K.ang_k;
        B(b, 3) = v.x; B(b, 4) = v.y; B(b, 5) = w;
    }

    // soft attraction force of agent b on the block (mrp00:421-424 / mrp02:470-474)
    MRP_HD V2 soft_force(int b, double force, int gb = 0) {
        double Ax = (double)B(b, 0), Ay = (double)B(b, 1), Bx = (double)B(gb, 0), By = (double)B(gb, 1);
        double dx = fabs(Bx - Ax), dy = fabs(By - Ay);
        double denom = dx > dy ? dx : dy;  // max(abs, abs)
        return mk((float)(force * ((Bx - Ax) / denom)), (float)(force * ((By - Ay) / denom)));
    }

    MRP_HD void control_v0(const float* a) {  // mrp00:415-424
        V2 bf = mk(0.0f, 0.0f);
        float btorque = 0.0f;
        V2 bc = mk(B(0, 0), B(0, 1));
        for (int i = 0; i < K.n; ++i) {
            int b = 1 + i;
            float x = a[3 * i], y = a[3 * i + 1], turn = a[3 * i + 2];
            B(b, 3) = (float)((double)x * K.SPEED);
            B(b, 4) = (float)((double)y * K.SPEED);
            B(b, 5) = turn;
            double force = pow(1.1, -gd(W_DIST + 2 * i));
            V2 f = soft_force(b, force);
            bf = bf + f;
            btorque += cross(bc - bc, f);
        }
        for (int i = 0; i < K.n; ++i) integrate_velocity(1 + i, mk(0.0f, 0.0f), 0.0f);
        integrate_velocity(0, bf, btorque);
    }

    MRP_HD void control_v2(const float* a) {  // mrp02:446-474
        V2 bf = mk(0.0f, 0.0f);
        float btorque = 0.0f;
        const int gb = goal_block();   // 0 unless the square variant has moved on to its next block
        V2 bc = mk(B(gb, 0), B(gb, 1));
        for (int i = 0; i < K.n; ++i) {
            int b = K.nblk + i;
            float turn = a[2 * i], vel = a[2 * i + 1];
            Xf xf = body_xf(b);
            V2 c = mk(B(b, 0), B(b, 1));
            V2 f = rmul(xf.q, mk(0.0f, 1.0f));   // GetWorldVector((0,1))
            V2 p = xmul(xf, mk(0.0f, 2.0f));     // GetWorldPoint((0,2))
            V2 ff = mk((float)((double)f.x * (double)vel * 0.75), (float)((double)f.y * (double)vel * 0.75));
            V2 force = ff;
            float torque = cross(p - c, ff);
            // updateFriction (mrp02:116-122)
            V2 rn = rmul(xf.q, mk(1.0f, 0.0f));
            V2 v = mk(B(b, 3), B(b, 4));
            float w = B(b, 5);
            float dn = dot(rn, v);
            V2 lat = dn * rn;
            V2 imp = K.ag_mass * (-lat);
            v = v + K.ag_invMass * imp;
            w += K.ag_invI * cross(c - c, imp);
            // ApplyAngularImpulse(0.1 * inertia * angularVelocity)
            w += K.ag_invI * (float)(0.1 * (double)K.ag_inertia * (double)w);
            B(b, 3) = v.x; B(b, 4) = v.y; B(b, 5) = w;
            double tq = (double)fabsf(turn) * 0.0005;
            float tturn = turn;
            if ((double)fabsf(vel) < 0.1) tturn = 0.0f;
            if (tturn < 0.0f) torque += (float)tq;
            else if (tturn > 0.0f) torque += (float)(-tq);
            else torque += 0.0f;
            double sf = pow(10.0, -gd(W_DIST + 2 * i));
            sf /= 50;
            V2 s = soft_force(b, sf, gb);
            bf = bf + s;
            btorque += cross(bc - bc, s);
            integrate_velocity(b, force, torque);
        }
        for (int k = 0; k < K.nblk; ++k) {
            if (k == gb) integrate_velocity(k, bf, btorque);
            else integrate_velocity(k, mk(0.0f, 0.0f), 0.0f);   // the other blocks only feel damping and contacts
        }
    }

    // distances, observation, reward, done after world.Step; returns env "done"
    MRP_HD bool post_step(float* obs, double* reward_out) {
        if (K.nblk > 1) return post_step_square(obs, reward_out);
        const int n = K.n;
        double prev_ad[MRP_MAX_AGENTS];
        for (int i = 0; i < n; ++i) prev_ad[i] = gd(W_DIST + 2 * i);
        double prev_bd = gd(W_DIST + 2 * n);
        double gx = gd(W_GOAL), gy = gd(W_GOAL + 2);
        double ad[MRP_MAX_AGENTS], bd;
        V2 bc = mk(B(0, 0), B(0, 1));
        if (!K.v2) {  // _calculate_distance / _calculate_agent_distance with b2Vec2*SCALE in float32
            float s = (float)K.SCALE;
            bd = py_distance((double)(bc.x * s), (double)(bc.y * s), gx, gy);
#pragma unroll 1
            for (int i = 0; i < n; ++i)
                ad[i] = py_distance((double)(B(1 + i, 0) * s), (double)(B(1 + i, 1) * s), (double)(bc.x * s), (double)(bc.y * s));
        } else {
            bd = py_distance((double)bc.x * K.ratio, (double)bc.y * K.ratio, gx, gy);
#pragma unroll 1
            for (int i = 0; i < n; ++i)
                ad[i] = py_distance((double)B(1 + i, 0) * K.ratio, (double)B(1 + i, 1) * K.ratio, (double)bc.x * K.ratio, (double)bc.y * K.ratio);
        }
#pragma unroll 1
        for (int i = 0; i < n; ++i) gsd(W_DIST + 2 * i, ad[i]);
        gsd(W_DIST + 2 * n, bd);

        Xf bxf = body_xf(0);
        obs_begin(obs);
        double reward = 0.0;
        bool in_place;
        int blks = (int)g(W_INPLACE);
        if (!K.v2) {
#pragma unroll 1
            for (int i = 0; i < n; ++i) {
                obs_put((float)((double)B(1 + i, 0) * K.SCALE - (double)bc.x * K.SCALE));
                obs_put((float)((double)B(1 + i, 1) * K.SCALE - (double)bc.y * K.SCALE));
                obs_put((float)ad[i]);
                obs_put(((goalc >> i) & 1) ? 1.0f : 0.0f);
            }
            double x = (double)bc.x * K.SCALE, y = (double)bc.y * K.SCALE;
            double angle = py_mod((double)B(0, 2), kTwoPiD);
            double a_diff = 0.0 - angle;
            in_place = !(fabs(gx - x) > 25.0) && !(fabs(gy - y) > 25.0);
            obs_put((float)(x - gx));
            obs_put((float)(y - gy));
            obs_put((float)a_diff);
            obs_put((float)py_distance(x, y, gx, gy));
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                V2 p = xmul(bxf, mk(K.blkv[k][0], K.blkv[k][1]));
                obs_put((float)((double)p.x * K.SCALE));
                obs_put((float)((double)p.y * K.SCALE));
            }
            obs_end();
            reward += (prev_bd - bd) * K.rp.blockDelta * 1.0 / 4.;
            reward -= K.rp.blockDistance * bd * 1.0 / 4.;
#pragma unroll 1
            for (int i = 0; i < n; ++i) {
                reward += (prev_ad[i] - ad[i]) * K.rp.agentDelta * 1.0 / 4.;
                reward -= K.rp.agentDistance * ad[i] * 1.0 / 4.;
                if ((goalc >> i) & 1) reward += 0.25;
            }
            int now = in_place ? 1 : 0;
            reward += (double)((now - blks) * 10);
            g(W_INPLACE) = (uint32_t)now;
            bool done = false;
            if (now == 1) { done = true; reward += 10000.0; }
            *reward_out = reward;
            return done;
        }
        // curriculum knobs: per-env vectors when installed (mrp_enable_curriculum), else the handle's scalars
        const int64_t env_ix = env_i;
        const double eps = K.eps_env ? K.eps_env[env_ix] : K.rp.scaled_epsilon;
        const double decay_pow = K.decay_env ? K.decay_env[env_ix] : K.rp.decay_pow;
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            int b = 1 + i;
            double aX = (double)B(b, 0) * K.ratio, aY = (double)B(b, 1) * K.ratio;
            double theta = py_mod((double)B(b, 2), kTwoPiD);
            double nt = theta <= kPiD ? -theta / kPiD : (kTwoPiD - theta) / kPiD;
            obs_put((float)aX);
            obs_put((float)aY);
            obs_put((float)nt);
            double bX = (double)bc.x * K.ratio, bY = (double)bc.y * K.ratio;
            obs_put((float)(aX - bX));
            obs_put((float)(aY - bY));
            obs_put(B(b, 3));
            obs_put(B(b, 4));
            obs_put(B(b, 5));
            obs_put((float)ad[i]);
        }
        {
            double x = (double)bc.x * K.ratio, y = (double)bc.y * K.ratio;
            double angle = py_mod((double)B(0, 2), kTwoPiD);
            double a_diff = (0.0 - angle) / kPiD;
            in_place = !(fabs(gx - x) > eps) && !(fabs(gy - y) > eps);
            obs_put((float)(x - gx));
            obs_put((float)(y - gy));
            obs_put((float)a_diff);
            obs_put((float)py_distance(x, y, gx, gy));
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                V2 p = xmul(bxf, mk(K.blkv[k][0], K.blkv[k][1]));
                obs_put((float)((double)p.x * K.ratio));
                obs_put((float)((double)p.y * K.ratio));
            }
        }
        obs_put((float)eps);
        obs_end();
        reward += (prev_bd - bd) * K.rp.blockDelta;
        reward -= K.rp.blockDistance * bd;
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            reward += (prev_ad[i] - ad[i]) * K.rp.agentDelta;
            reward -= K.rp.agentDistance * ad[i];
        }
        const double BOUNDS = 0.1;
        bool agt_oob = false;
#pragma unroll 1
        for (int i = 0; i < n && !agt_oob; ++i) {
            double x = (double)B(1 + i, 0), y = (double)B(1 + i, 1);
            if (x < BOUNDS || x > (K.W - BOUNDS)) agt_oob = true;
            else if (y < BOUNDS || y > (K.H - BOUNDS)) agt_oob = true;
        }
        if (agt_oob) {
            reward -= K.rp.outOfBounds * decay_pow;
            *reward_out = reward;
            return true;
        }
        {
            double x = (double)bc.x, y = (double)bc.y;
            bool oob = (x < BOUNDS || x > (K.W - BOUNDS)) || (y < BOUNDS || y > (K.H - BOUNDS));
            if (oob) {
                reward -= K.rp.blkOutOfBounds * decay_pow;
                *reward_out = reward;
                return true;
            }
        }
        int now = in_place ? 1 : 0;
        g(W_INPLACE) = (uint32_t)now;
        int num_in_contact = 0;
#pragma unroll 1
        for (int i = 0; i < n; ++i) num_in_contact += (goalc >> i) & 1;
        bool done = false;
        if (now == 1) {
            done = true;
            reward += K.rp.puzzleComp * decay_pow * ((double)num_in_contact / (double)n);
        }
        *reward_out = reward;
        return done;
    }

    // Square variant (BASELINE.json configs[4]; semantics: DESIGN.md "Square variant", modelled on mrp02:491-584):
    // per robot the 9 values of v2 relative to the goal block; per block (x - tx, y - ty, (ta - angle) / pi, distance to its
    // target, vertices); epsilon, goal-block index, blocks in place, robots touching the goal block.  When the goal block reaches
    // its target the next block of the queue T, L, I becomes the goal; done when all three are placed.
    MRP_HD bool post_step_square(float* obs, double* reward_out) {
        const int n = K.n, nbk = K.nblk;
        const int placed = (int)g(W_INPLACE);
        const int gb = placed < nbk - 1 ? placed : nbk - 1;
        double prev_ad[MRP_MAX_AGENTS], ad[MRP_MAX_AGENTS];
        for (int i = 0; i < n; ++i) prev_ad[i] = gd(W_DIST + 2 * i);
        const double prev_bd = gd(W_DIST + 2 * n);
        const double gx = gd(W_GOAL), gy = gd(W_GOAL + 2);
        const V2 bc = mk(B(gb, 0), B(gb, 1));
        const double bd = py_distance((double)bc.x * K.ratio, (double)bc.y * K.ratio, gx + K.sq_target[gb][0] * K.ratio, gy + K.sq_target[gb][1] * K.ratio);
#pragma unroll 1
        for (int i = 0; i < n; ++i)
            ad[i] = py_distance((double)B(nbk + i, 0) * K.ratio, (double)B(nbk + i, 1) * K.ratio, (double)bc.x * K.ratio, (double)bc.y * K.ratio);
        const double eps = K.eps_env ? K.eps_env[env_i] : K.rp.scaled_epsilon;
        const double decay_pow = K.decay_env ? K.decay_env[env_i] : K.rp.decay_pow;
        obs_begin(obs);
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const int b = nbk + i;
            const double aX = (double)B(b, 0) * K.ratio, aY = (double)B(b, 1) * K.ratio;
            const double theta = py_mod((double)B(b, 2), kTwoPiD);
            const double nt = theta <= kPiD ? -theta / kPiD : (kTwoPiD - theta) / kPiD;
            obs_put((float)aX);
            obs_put((float)aY);
            obs_put((float)nt);
            const double bX = (double)bc.x * K.ratio, bY = (double)bc.y * K.ratio;
            obs_put((float)(aX - bX));
            obs_put((float)(aY - bY));
            obs_put(B(b, 3));
            obs_put(B(b, 4));
            obs_put(B(b, 5));
            obs_put((float)ad[i]);
        }
        bool in_place = false;
#pragma unroll 1
        for (int k = 0; k < nbk; ++k) {
            const Xf bxf = body_xf(k);
            const double x = (double)B(k, 0) * K.ratio, y = (double)B(k, 1) * K.ratio;
            const double angle = py_mod((double)B(k, 2), kTwoPiD);
            const double fx = gx + K.sq_target[k][0] * K.ratio, fy = gy + K.sq_target[k][1] * K.ratio;
            const double a_diff = (K.sq_target[k][2] - angle) / kPiD;
            if (k == gb) in_place = !(fabs(fx - x) > eps) && !(fabs(fy - y) > eps);
            obs_put((float)(x - fx));
            obs_put((float)(y - fy));
            obs_put((float)a_diff);
            obs_put((float)py_distance(x, y, fx, fy));
#pragma unroll 1
            for (int v = K.blkv_off[k]; v < K.blkv_off[k + 1]; ++v) {
                const V2 p = xmul(bxf, mk(K.blkv[v][0], K.blkv[v][1]));
                obs_put((float)((double)p.x * K.ratio));
                obs_put((float)((double)p.y * K.ratio));
            }
        }
        int touching = 0;
#pragma unroll 1
        for (int i = 0; i < n; ++i) touching += (goalc >> i) & 1;
        obs_put((float)eps);
        obs_put((float)gb);
        obs_put((float)placed);
        obs_put((float)touching);
        obs_end();
        double reward = 0.0;
        reward += (prev_bd - bd) * K.rp.blockDelta;
        reward -= K.rp.blockDistance * bd;
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            reward += (prev_ad[i] - ad[i]) * K.rp.agentDelta;
            reward -= K.rp.agentDistance * ad[i];
        }
#pragma unroll 1
        for (int i = 0; i < n; ++i) gsd(W_DIST + 2 * i, ad[i]);
        gsd(W_DIST + 2 * n, bd);
        const double BOUNDS = 0.1;
#pragma unroll 1
        for (int b = nbk; b < K.nb; ++b) {   // robots first, then the blocks (the order decides which penalty applies)
            const double x = (double)B(b, 0), y = (double)B(b, 1);
            if ((x < BOUNDS || x > (K.W - BOUNDS)) || (y < BOUNDS || y > (K.H - BOUNDS))) {
                *reward_out = reward - K.rp.outOfBounds * decay_pow;
                return true;
            }
        }
#pragma unroll 1
        for (int b = 0; b < nbk; ++b) {
            const double x = (double)B(b, 0), y = (double)B(b, 1);
            if ((x < BOUNDS || x > (K.W - BOUNDS)) || (y < BOUNDS || y > (K.H - BOUNDS))) {
                *reward_out = reward - K.rp.blkOutOfBounds * decay_pow;
                return true;
            }
        }
        bool done = false;
        if (in_place) {
            reward += K.rp.puzzleComp * decay_pow * ((double)touching / (double)n);
            g(W_INPLACE) = (uint32_t)(placed + 1);
            if (placed + 1 == nbk) done = true;
            else {   // the next block of the queue becomes the goal: distances re-based on it, contact flags start over
                const int g2 = placed + 1;
                goalc = 0;
                g(W_GOALC) = 0;   // store() may already have run (staged observation rows)
                const V2 c2 = mk(B(g2, 0), B(g2, 1));
                gsd(W_DIST + 2 * n, py_distance((double)c2.x * K.ratio, (double)c2.y * K.ratio, gx + K.sq_target[g2][0] * K.ratio,
                                                gy + K.sq_target[g2][1] * K.ratio));
#pragma unroll 1
                for (int i = 0; i < n; ++i)
                    gsd(W_DIST + 2 * i, py_distance((double)B(nbk + i, 0) * K.ratio, (double)B(nbk + i, 1) * K.ratio, (double)c2.x * K.ratio,
                                                    (double)c2.y * K.ratio));
            }
        }
        *reward_out = reward;
        return done;
    }

    // one env.step body: control -> world.Step -> post
    MRP_HD bool env_step(const float* a, float* obs, double* reward, bool new_fixtures) {
        if (K.v2) control_v2(a); else control_v0(a);
        world_step(new_fixtures);
        return post_step(obs, reward);
    }

    // respawn (mrp00:392-409 / mrp02:421-440) with counter-based Philox (north star)
    MRP_HD void spawn(uint32_t episode) {
        uint32_t d = 0;
        const double W = K.W, H = K.H;
        double blk[3][3];   // pose of every block
        double ag[2 * MRP_MAX_AGENTS];
        if (K.nblk > 1) {
            // square variant: the blocks start on the vertical centre line, each with its own random angle; robots and goal as v2
            g(W_INPLACE) = 0;   // the block queue starts over
            for (int k = 0; k < K.nblk; ++k) {
                blk[k][0] = W / 2; blk[k][1] = H * (k + 1) / 4;
                blk[k][2] = 0.0 + (kTwoPiD - 0.0) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
            }
            for (int i = 0; i < K.n; ++i) {
                ag[2 * i] = 0.3 + ((W / 3 - 0.3) - 0.3) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
                ag[2 * i + 1] = 0.3 + ((H - 0.3) - 0.3) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
            }
        } else if (!K.v2) {
            blk[0][0] = 1.0 + ((W - 1.0) - 1.0) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
            blk[0][1] = 1.0 + ((H - 1.0) - 1.0) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
            blk[0][2] = 0.0 + (kTwoPiD - 0.0) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
            for (int i = 0; i < K.n; ++i) {
                ag[2 * i] = 1.0 + ((W - 1.0) - 1.0) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
                ag[2 * i + 1] = 1.0 + ((H - 1.0) - 1.0) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
            }
        } else {
            blk[0][0] = W / 2; blk[0][1] = H / 2;
            blk[0][2] = 0.0 + (kTwoPiD - 0.0) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
            for (int i = 0; i < K.n; ++i) {
                ag[2 * i] = 0.3 + ((W / 3 - 0.3) - 0.3) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
                ag[2 * i + 1] = 0.3 + ((H - 0.3) - 0.3) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
            }
        }
        // bodies: xf.p = position, sweep.c = Mul(xf, localCenter) (b2Body ctor + ResetMassData)
        for (int b = 0; b < K.nb; ++b) {
            const bool isb = b < K.nblk;
            float px = (float)(isb ? blk[b][0] : ag[2 * (b - K.nblk)]);
            float py = (float)(isb ? blk[b][1] : ag[2 * (b - K.nblk) + 1]);
            float ang = isb ? (float)blk[b][2] : (K.v2 ? (float)(3.0 / 2 * kPiD) : 0.0f);
            Xf xf;
            xf.p = mk(px, py);
            xf.q = rot_set(ang);
            V2 c = xmul(xf, localCenter(b));
            B(b, 0) = c.x; B(b, 1) = c.y; B(b, 2) = ang;
            B(b, 3) = 0.0f; B(b, 4) = 0.0f; B(b, 5) = 0.0f;
            BX(b, 6) = xf.q.s; BX(b, 7) = xf.q.c; BX(b, 8) = px; BX(b, 9) = py;
            set_rot_cache(b, xf.q, ang);
            // proxies: fat AABB = tight AABB at creation +- 0.1 (b2DynamicTree::CreateProxy)
            int f0, f1;
            fix_range(b, f0, f1);
            for (int f = f0; f < f1; ++f) {
                Box t = shape_aabb(fix_shape(f), xf);
                FA(f, 0) = t.lx - kAabbExtension; FA(f, 1) = t.ly - kAabbExtension;
                FA(f, 2) = t.hx + kAabbExtension; FA(f, 3) = t.hy + kAabbExtension;
            }
        }
        for (int k = 0; k < 4; ++k) {
            int b = K.nb + k;
            B(b, 0) = ct[CT_WALLPOS + 2 * k];
            B(b, 1) = ct[CT_WALLPOS + 2 * k + 1];
            B(b, 2) = 0.0f; B(b, 3) = 0.0f; B(b, 4) = 0.0f; B(b, 5) = 0.0f;
        }
        nc = 0;
        goalc = 0;
        double gx = K.goal_x0, gy = K.goal_y0;
        if (K.v2) {  // _set_random_goal (mrp02:303-311)
            double x = (W * 2 / 3 + 0.4) + ((W - 0.4) - (W * 2 / 3 + 0.4)) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
            double y = 0.4 + ((H - 0.4) - 0.4) * uniform53(K.seed, kStreamSpawn, gid, episode, d++);
            gx = x * K.ratio;
            gy = y * K.ratio;
        }
        gsd(W_GOAL, gx);
        gsd(W_GOAL + 2, gy);
        // _calculate_distance / _calculate_agent_distance
        V2 bc = mk(B(0, 0), B(0, 1));
        if (!K.v2) {
            float s = (float)K.SCALE;
            gsd(W_DIST + 2 * K.n, py_distance((double)(bc.x * s), (double)(bc.y * s), gx, gy));
            for (int i = 0; i < K.n; ++i)
                gsd(W_DIST + 2 * i, py_distance((double)(B(1 + i, 0) * s), (double)(B(1 + i, 1) * s), (double)(bc.x * s), (double)(bc.y * s)));
        } else {
            // square variant: the goal block is block 0 and its target sits at an offset from the goal centre
            const double tx = K.nblk > 1 ? gx + K.sq_target[0][0] * K.ratio : gx, ty = K.nblk > 1 ? gy + K.sq_target[0][1] * K.ratio : gy;
            gsd(W_DIST + 2 * K.n, py_distance((double)bc.x * K.ratio, (double)bc.y * K.ratio, tx, ty));
            for (int i = 0; i < K.n; ++i)
                gsd(W_DIST + 2 * i, py_distance((double)B(K.nblk + i, 0) * K.ratio, (double)B(K.nblk + i, 1) * K.ratio, (double)bc.x * K.ratio, (double)bc.y * K.ratio));
        }
    }

    // ---------------------------------------------------------------- phase pipeline (one env.step split at the solver)
    // k_pre: control, Collide, island order; envs with touching contacts become solver tasks, the others
    // integrate their positions right away.  Returns T.
    MRP_HD int pre_phase(const float* a, unsigned walls_pending = 0u) {
        // walls_pending (lanes of the warp that run this pass): the warp's action rows were staged in the wall slots; once every
        // lane has its row in registers the slots get their walls
        if (walls_pending) {
#if defined(__CUDA_ARCH__)
            __syncwarp(walls_pending);
#endif
            init_walls();
        }
        if (K.v2) control_v2(a); else control_v0(a);
        finish_collide();
        // pose at the start of the step, for k_post's SynchronizeFixtures / TOI sweeps
        for (int b = 0; b < K.nb; ++b) {
            const int w = K.w_body + kBodyWords * b;
            gsf(w + 8, B(b, 0)); gsf(w + 9, B(b, 1)); gsf(w + 10, B(b, 2));
        }
        uint8_t island_of[kMaxC];
        const int T = build_islands(island_of);
        uint32_t in_island = 0;  // dynamic bodies that belong to an island with touching contacts
        if (T > 0) {
            const int off = atomic_add_i32(&K.cnt[CNT_POOL], T);  // in records
            MRP_ASSERT(off >= 0 && (int64_t)off + T <= (int64_t)K.nloc * K.maxc, CHK_RECORD);
            vcp = K.pool + (size_t)off * VC_WORDS;
            init_constraints(T, island_of, true);
            warm_start(T);
            for (int t = 0; t < T; ++t) { if (((vmeta(t) >> 8) & 3) == 2) ++stat_m2; else ++stat_m1; }
            // one solver task per island (its constraint records are contiguous in solver order)
            const bool heavy = g(W_HINT) >= kHeavyHint;
            const int cap = K.nloc * K.nb;
            for (int start = 0; start < T;) {
                int end = start + 1;
                while (end < T && island_of[end] == island_of[start]) ++end;
                const int cls = end - start > 2 ? 3 : (end - start == 2 ? 2 : (int)((vmeta(start) >> 8) & 3) - 1);  // single contact: vpc is 1 or 2
                // two-contact islands are queued by manifold kind instead: 88 % of them pair two 1-point manifolds (settled Heavy-v0
                // rollout), and a single lane with a 2-point manifold makes its warp run the second friction point and the block
                // solver as well — 206 instead of 84 instructions per contact and sweep.  The ones with a 2-point manifold go to the
                // front of the queue (they are the longer solves too), the pure 1-point pairs to the back, so warps stay uniform
                const bool front = cls == 2 ? (((vmeta(start) >> 8) & 3) == 2 || ((vmeta(start + 1) >> 8) & 3) == 2) : heavy;
                const int task = cls * cap + (front ? atomic_add_i32(&K.cnt[CNT_TASKS + cls], 1)
                                                    : cap - 1 - atomic_add_i32(&K.cnt[CNT_TASKS_LIGHT + cls], 1));
                MRP_ASSERT(task >= 0 && task < kTaskClasses * cap, CHK_QUEUE);
                K.task_env[task] = env_i;
                K.task_T[task] = end - start;
                K.task_off[task] = off + start;
                start = end;
            }
            for (int t = 0; t < T; ++t) {
                const uint32_t m = meta[order[t]];
                in_island |= (1u << ((m >> 20) & 15)) | (1u << ((m >> 24) & 15));
            }
        }
        g(W_HINT) = 0;  // solver tasks of this env add their operation counts
        for (int b = 0; b < K.nb; ++b)
            if (!((in_island >> b) & 1)) integrate_position(b, K.h);
        // hand-off to the solver / post kernels: poses + velocities, pre-step pose, contact list
        g(W_NC) = (uint32_t)nc;
        g(W_GOALC) = goalc;
        for (int b = 0; b < K.nb; ++b) {
            const int w = K.w_body + kBodyWords * b;
            for (int f = 0; f < 6; ++f) gsf(w + f, B(b, f));
        }
        for (int k = 0; k < nc; ++k) g(cw(k, 0)) = meta[k];
        return T;
    }
    // k_post: the step after the solver — transforms, broadphase, TOI, obs / reward / done.
    // Returns false (nothing stored) when a TOI event is needed and allow_events is false.
    // entry != 0 (k_post on the device): the lanes `entry` of the warp run this pass, and the observation rows leave through the
    // warp's staging block (obs_stage / obs_flush) instead of row by row
    MRP_HD bool post_phase(float* obs, double* reward, bool* done_env, bool allow_events, unsigned entry = 0u) {
        // load(): q/p are still the pre-step transform; c0/a0 from the hand-off words
        for (int b = 0; b < K.nb; ++b) set_rot_cache(b, Rot{BX(b, 6), BX(b, 7)}, BX(b, c0f + 2));
        const bool fin = post_solve(allow_events);
#if defined(__CUDA_ARCH__)
        if (entry) obs_stage(entry, fin);
#else
        (void)entry;
#endif
        if (!fin) return false;
        *done_env = post_step(obs, reward);
        return true;
    }

    // env.reset(): respawn + hidden step with a sampled action (mrp00:411 / mrp02:442)
    MRP_HD void reset_env(float* obs) {
        uint32_t episode = g(W_EPISODE) + 1u;
        g(W_EPISODE) = episode;
        g(W_ELAPSED) = 0;
        g(W_EPLEN) = 0;
        gsd(W_EPRET, 0.0);
        spawn(episode);
        float a[3 * MRP_MAX_AGENTS];
        for (int i = 0; i < K.act_dim; ++i)
            a[i] = (float)(-1.0 + 2.0 * uniform53(K.seed, kStreamResetAction, gid, episode, (uint32_t)i));
        double r;
        env_step(a, obs, &r, true);
        store();
    }
};

}  // namespace mrp
