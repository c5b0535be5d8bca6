// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE ONLY.
// C entry points (ctypes) around orc::Env for tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs.  Nothing in gym_puzzles_b200/
// may load this library.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "mrp_env.hpp"

using orc::Env;

struct OrcBatch {
    std::vector<Env*> envs;
    mrp_layout L;
    int nthreads;
};

template <class F>
static void parallel_for(int n, int nthreads, F f) {
    if (nthreads <= 1 || n < 2) {
        for (int i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<int> next(0);
    std::vector<std::thread> ts;
    int nt = std::min(nthreads, n);
    for (int t = 0; t < nt; ++t)
        ts.emplace_back([&]() {
            for (;;) {
                int i = next.fetch_add(16);
                if (i >= n) break;
                int e = std::min(n, i + 16);
                for (int k = i; k < e; ++k) f(k);
            }
        });
    for (auto& t : ts) t.join();
}

extern "C" {

void* orc_create(int variant, int n_agents, int num_envs, uint64_t seed, uint64_t env_id_base, int nthreads) {
    mrp_layout L;
    if (mrp_layout_for(variant, n_agents, &L) != 0 || num_envs <= 0) return nullptr;
    OrcBatch* b = new OrcBatch();
    b->L = L;
    b->nthreads = nthreads > 0 ? nthreads : 1;
    b->envs.resize(num_envs);
    for (int i = 0; i < num_envs; ++i) b->envs[i] = new Env(variant, L.n_agents, seed, env_id_base + (uint64_t)i);
    return b;
}
void orc_destroy(void* h) {
    OrcBatch* b = (OrcBatch*)h;
    if (!b) return;
    for (Env* e : b->envs) delete e;
    delete b;
}
int orc_layout(void* h, mrp_layout* out) { *out = ((OrcBatch*)h)->L; return 0; }

int orc_set_auto_reset(void* h, int on) {
    for (Env* e : ((OrcBatch*)h)->envs) e->auto_reset = on != 0;
    return 0;
}
int orc_set_max_episode_steps(void* h, int steps) {
    OrcBatch* b = (OrcBatch*)h;
    b->L.max_episode_steps = steps;
    for (Env* e : b->envs) e->L.max_episode_steps = steps;
    return 0;
}
// reward weights (set_reward_params mrp00:231-239), epsilon (update_goal mrp02:232-233), decay^(-t) (update_params mrp02:227-230)
int orc_set_params(void* h, const double* p9) {
    for (Env* e : ((OrcBatch*)h)->envs) {
        e->rp.agentDelta = p9[0]; e->rp.agentDistance = p9[1]; e->rp.blockDelta = p9[2]; e->rp.blockDistance = p9[3];
        e->rp.puzzleComp = p9[4]; e->rp.outOfBounds = p9[5]; e->rp.blkOutOfBounds = p9[6];
        e->rp.scaled_epsilon = p9[7]; e->rp.decay_pow = p9[8];
    }
    return 0;
}
// per-env curriculum vectors (update_goal / update_params per env; checker for mrp_enable_curriculum)
int orc_set_curriculum(void* h, const double* eps, const double* decay_pow) {
    OrcBatch* b = (OrcBatch*)h;
    for (size_t i = 0; i < b->envs.size(); ++i) {
        if (eps) b->envs[i]->rp.scaled_epsilon = eps[i];
        if (decay_pow) b->envs[i]->rp.decay_pow = decay_pow[i];
    }
    return 0;
}
int orc_get_params(void* h, double* p9) {
    Env* e = ((OrcBatch*)h)->envs[0];
    p9[0] = e->rp.agentDelta; p9[1] = e->rp.agentDistance; p9[2] = e->rp.blockDelta; p9[3] = e->rp.blockDistance;
    p9[4] = e->rp.puzzleComp; p9[5] = e->rp.outOfBounds; p9[6] = e->rp.blkOutOfBounds;
    p9[7] = e->rp.scaled_epsilon; p9[8] = e->rp.decay_pow;
    return 0;
}

// obs: double[num_envs][obs_dim]; mask may be NULL (all)
int orc_reset(void* h, const uint8_t* mask, double* obs) {
    OrcBatch* b = (OrcBatch*)h;
    int O = b->L.obs_dim;
    parallel_for((int)b->envs.size(), b->nthreads, [&](int i) {
        if (mask && !mask[i]) return;
        std::vector<double> tmp(O);
        b->envs[i]->reset(obs ? obs + (size_t)i * O : tmp.data());
    });
    return 0;
}
int orc_step(void* h, const float* actions, double* obs, double* reward, uint8_t* done, uint8_t* trunc) {
    OrcBatch* b = (OrcBatch*)h;
    int O = b->L.obs_dim, A = b->L.act_dim;
    parallel_for((int)b->envs.size(), b->nthreads, [&](int i) {
        b->envs[i]->step(actions + (size_t)i * A, obs + (size_t)i * O, reward + i, done + i, trunc + i);
    });
    return 0;
}
// synthetic action stream of the benchmark: U(-1,1) f32, Philox stream ACTION, counter (env, step)
int orc_sample_actions(void* h, uint64_t step_index, float* actions) {
    OrcBatch* b = (OrcBatch*)h;
    int A = b->L.act_dim;
    for (size_t i = 0; i < b->envs.size(); ++i) {
        Env* e = b->envs[i];
        for (int k = 0; k < A; ++k)
            actions[i * A + k] = (float)(-1.0 + 2.0 * orc::uniform53(e->seed, orc::kStreamAction, e->gid, (uint32_t)step_index, (uint32_t)k));
    }
    return 0;
}
int orc_get_state(void* h, int env_begin, int env_count, uint32_t* words) {
    OrcBatch* b = (OrcBatch*)h;
    for (int i = 0; i < env_count; ++i) b->envs[env_begin + i]->get_state(words + (size_t)i * b->L.state_words);
    return 0;
}
int orc_set_state(void* h, int env_begin, int env_count, const uint32_t* words) {
    OrcBatch* b = (OrcBatch*)h;
    parallel_for(env_count, b->nthreads, [&](int i) { b->envs[env_begin + i]->set_state(words + (size_t)i * b->L.state_words); });
    return 0;
}
// {episodes, successes(done by env), truncations, sum_return, sum_len, toi_events, toi_calls, pos_iters}
int orc_stats(void* h, double* out8) {
    OrcBatch* b = (OrcBatch*)h;
    for (int k = 0; k < 8; ++k) out8[k] = 0;
    for (Env* e : b->envs) {
        out8[0] += e->n_episodes; out8[1] += e->n_success; out8[2] += e->n_trunc;
        out8[3] += e->sum_return; out8[4] += e->sum_len;
        out8[5] += e->world->stat_toi_events; out8[6] += e->world->stat_toi_calls; out8[7] += e->world->stat_pos_iters;
    }
    return 0;
}

// workload probe: enable, then read {steps, contacts, touching, islands_with_contacts, fix_iters, fix_never, hist[8]}
int orc_probe(void* h, int enable, double* out14, double* period9) {
    OrcBatch* b = (OrcBatch*)h;
    if (out14) for (int k = 0; k < 14; ++k) out14[k] = 0;
    if (period9) for (int k = 0; k < 9; ++k) period9[k] = 0;
    for (Env* e : b->envs) {
        b2o::World* w = e->world;
        if (out14) {
            out14[0] += w->stat_steps; out14[1] += w->stat_contact_sum; out14[2] += w->stat_touch_sum;
            out14[3] += w->stat_islands; out14[4] += w->stat_fix_iters; out14[5] += w->stat_fix_never;
            for (int k = 0; k < 8; ++k) out14[6 + k] += w->stat_islands_c[k];
        }
        if (period9) for (int k = 0; k < 9; ++k) period9[k] += w->stat_period[k];
        w->probe = enable != 0;
    }
    return 0;
}

// position-solver / long-cycle part of the workload probe: {pos islands, pos islands using all iterations, of those: cycle of
// period 1, 2, 3..8 found, sum of the iteration it was found at, iterations of all islands, sum of closing sweeps of velocity cycles}
int orc_probe_pos(void* h, double* out8) {
    OrcBatch* b = (OrcBatch*)h;
    for (int k = 0; k < 8; ++k) out8[k] = 0;
    for (Env* e : b->envs) {
        b2o::World* w = e->world;
        out8[0] += w->stat_pos_islands; out8[1] += w->stat_pos_full;
        out8[2] += w->stat_pos_cyc[0]; out8[3] += w->stat_pos_cyc[1]; out8[4] += w->stat_pos_cyc[2];
        out8[5] += w->stat_pos_cyc_at; out8[6] += w->stat_pos_iter_sum; out8[7] += w->stat_vel_cyc_at;
    }
    return 0;
}

// Box2D version fork of the collision routine (b2core.hpp g_fork_230): 0 = >= 2.3.1 (default), 1 = 2.3.0.  Process-wide.
int orc_set_box2d_fork(int fork_230) {
    int old = b2o::g_fork_230;
    b2o::g_fork_230 = fork_230 ? 1 : 0;
    return old;
}

// ---- primitive-level probes for known-answer tests ---------------------------------
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out4) {
    orc::Philox4 r = orc::philox4x32_10(c0, c1, c2, c3, k0, k1);
    std::memcpy(out4, r.v, 16);
}
// mass data of body b of env 0: {mass, invMass, I, invI, localCenter.x, localCenter.y}
void orc_body_mass(void* h, int body, float* out6) {
    Env* e = ((OrcBatch*)h)->envs[0];
    const b2o::Body& B = e->world->bodies[body];
    out6[0] = B.mass; out6[1] = B.invMass; out6[2] = B.I; out6[3] = B.invI;
    out6[4] = B.sweep.localCenter.x; out6[5] = B.sweep.localCenter.y;
}
// fixture f of env 0: count, then verts (x,y)*8, normals (x,y)*8, friction
int orc_fixture(void* h, int f, float* out33) {
    Env* e = ((OrcBatch*)h)->envs[0];
    const b2o::Fixture& F = e->world->fixtures[f];
    for (int i = 0; i < 8; ++i) {
        out33[2 * i] = F.shape.v[i].x; out33[2 * i + 1] = F.shape.v[i].y;
        out33[16 + 2 * i] = F.shape.n[i].x; out33[16 + 2 * i + 1] = F.shape.n[i].y;
    }
    out33[32] = F.friction;
    return F.shape.count;
}
// polygon-vs-polygon manifold of two boxes/polys given explicitly (KATs)
// polyA/polyB: count then verts; xf: {p.x,p.y,angle}; out: {pointCount,type,ln.x,ln.y,lp.x,lp.y, p0.x,p0.y,key0, p1.x,p1.y,key1}
void orc_collide(int nA, const float* vA, const float* xfA3, int nB, const float* vB, const float* xfB3, float* out12) {
    b2o::Polygon A, B;
    b2o::Vec2 pa[8], pb[8];
    for (int i = 0; i < nA; ++i) pa[i] = b2o::Vec2(vA[2 * i], vA[2 * i + 1]);
    for (int i = 0; i < nB; ++i) pb[i] = b2o::Vec2(vB[2 * i], vB[2 * i + 1]);
    A.Set(pa, nA);
    B.Set(pb, nB);
    b2o::Transform ta, tb;
    ta.p = b2o::Vec2(xfA3[0], xfA3[1]); ta.q.Set(xfA3[2]);
    tb.p = b2o::Vec2(xfB3[0], xfB3[1]); tb.q.Set(xfB3[2]);
    b2o::Manifold m;
    b2o::CollidePolygons(&m, &A, ta, &B, tb);
    out12[0] = (float)m.pointCount; out12[1] = (float)m.type;
    out12[2] = m.localNormal.x; out12[3] = m.localNormal.y; out12[4] = m.localPoint.x; out12[5] = m.localPoint.y;
    for (int j = 0; j < 2; ++j) {
        out12[6 + 3 * j] = m.points[j].localPoint.x; out12[7 + 3 * j] = m.points[j].localPoint.y;
        out12[8 + 3 * j] = (float)m.points[j].id.key();
    }
}
// TOI of two polygons with linear sweeps: sweep = {c0.x,c0.y,a0,c.x,c.y,a}; returns state, *t
int orc_toi(int nA, const float* vA, const float* sA6, int nB, const float* vB, const float* sB6, float* t) {
    b2o::Polygon A, B;
    b2o::Vec2 pa[8], pb[8];
    for (int i = 0; i < nA; ++i) pa[i] = b2o::Vec2(vA[2 * i], vA[2 * i + 1]);
    for (int i = 0; i < nB; ++i) pb[i] = b2o::Vec2(vB[2 * i], vB[2 * i + 1]);
    A.Set(pa, nA);
    B.Set(pb, nB);
    b2o::DistanceProxy dA, dB;
    dA.Set(&A); dB.Set(&B);
    b2o::Sweep a, b;
    a.c0 = b2o::Vec2(sA6[0], sA6[1]); a.a0 = sA6[2]; a.c = b2o::Vec2(sA6[3], sA6[4]); a.a = sA6[5];
    b.c0 = b2o::Vec2(sB6[0], sB6[1]); b.a0 = sB6[2]; b.c = b2o::Vec2(sB6[3], sB6[4]); b.a = sB6[5];
    return (int)b2o::TimeOfImpact(t, &dA, &dB, a, b, 1.0f);
}

}  // extern "C"
