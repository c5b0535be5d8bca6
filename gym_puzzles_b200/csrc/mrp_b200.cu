// mrp_b200.cu — sm_100a kernels + the C-ABI of include/mrp_b200.h.
//
// Build (see __graft_entry__.build()):
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -shared ...
// -fmad=false is load-bearing: Box2D on x86-64 never fuses mul+add, and contact / done
// flags must be bit-exact (BASELINE.json north_star).
//
// The same translation unit also builds as plain C++ with -DMRP_HOST_EMU (g++ -x c++) into
// tests/emu/libmrp_emu.so: a host execution of the *kernel source* used only to debug
// kernel logic in the GPU-less build container.  The Python package never loads it.
//
// This file is compiled TWICE into the library: once as is (contact capacity 32: every registered variant) and once
// with -DMRP_MAXC=192 ("wide": MultiRobotPuzzle2(num_agents > 2), mrp02:139).  The wide compilation renames its
// namespace, its handle type and its entry points (suffix _wide) so the two live side by side; the default build's
// entry points forward to the wide ones for handles created with such a configuration.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <initializer_list>
#include <new>

#ifndef MRP_MAXC
#define MRP_MAXC 32
#endif
#if MRP_MAXC > 32
#define MRP_WIDE 1
#define mrp mrp_wide                 // namespace of all inline code (the two builds differ in array sizes: keep the ODR)
#define mrp_handle mrp_handle_wide
#define MRP_API(name) name##_wide
#else
#define MRP_API(name) name
#endif

#include "mrp_env.cuh"
#include "mrp_variant.hpp"

#ifndef MRP_HOST_EMU
#include <cuda_runtime.h>
#endif

using namespace mrp;

// ------------------------------------------------------------------------------------
// kernels (device) / loops (host emulation)
// ------------------------------------------------------------------------------------
namespace {

constexpr int kCtPad = (CT_WORDS + 31) & ~31;

MRP_HD void stat_add(double* stats, int slot, double v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(stats + slot, v);
#else
    stats[slot] += v;
#endif
}

// workload counters: one atomic per warp (every lane of the warp calls)
#ifndef MRP_HOST_EMU
__device__ __forceinline__ void stat_add_warp(double* stats, int slot, uint32_t v) {
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(stats + slot, (double)v);
}
#endif

// TimeLimit, episode accounting, auto-reset queue: the tail of every env.step
MRP_HD void finish_step(const SimConst& K, Env& e, int64_t env, bool d, double r) {
    uint32_t elapsed = e.g(W_ELAPSED) + 1u;
    e.g(W_ELAPSED) = elapsed;
    // NaN / inf guard: a dynamic body whose pose or velocity is not finite ends the episode by force —
    // reported like a TimeLimit truncation with reward 0, counted in MRP_STAT_NAN_RESETS, respawned by the auto-reset
    uint32_t expo = 0u;
    for (int b = 0; b < K.nb; ++b)
        for (int f = 0; f < 6; ++f) expo |= ((Sim::__float_as_uint_(e.B(b, f)) & 0x7f800000u) == 0x7f800000u) ? 1u : 0u;
    bool limit = (int)elapsed >= K.max_steps;
    if (expo) {
        d = false; limit = true; r = 0.0;
        stat_add(K.stats, MRP_STAT_NAN_RESETS, 1.0);
    }
    bool done = d || limit;
    K.rew[env] = (float)r;
    K.done[env] = done ? 1 : 0;
    K.trunc[env] = (limit && !d) ? 1 : 0;
    double ret = e.gd(W_EPRET) + r;
    uint32_t len = e.g(W_EPLEN) + 1u;
    e.gsd(W_EPRET, ret);
    e.g(W_EPLEN) = len;
    if (!e.stored) e.store();
    if (e.overflow) stat_add(K.stats, MRP_STAT_OVERFLOW, 1.0);
    if (done) {
        if (K.term_obs) {  // keep what the auto-reset is about to overwrite
            const float* orow = K.obs + env * K.obs_dim;
            float* trow = K.term_obs + env * K.obs_dim;
#if defined(__CUDA_ARCH__)
            for (int i = 0; i < K.obs_dim; ++i) trow[i] = __ldcg(orow + i);   // the row may have been written by other lanes of the warp (obs_flush)
#else
            for (int i = 0; i < K.obs_dim; ++i) trow[i] = orow[i];
#endif
            K.term_ret[env] = (float)ret;
            K.term_len[env] = (int32_t)len;
        }
        stat_add(K.stats, MRP_STAT_EPISODES, 1.0);
        if (d) stat_add(K.stats, MRP_STAT_DONE_BY_ENV, 1.0);
        if (limit && !d) stat_add(K.stats, MRP_STAT_TRUNCATED, 1.0);
        stat_add(K.stats, MRP_STAT_SUM_RETURN, ret);
        stat_add(K.stats, MRP_STAT_SUM_RETURN_SQ, ret * ret);
        stat_add(K.stats, MRP_STAT_SUM_LENGTH, (double)len);
        if (K.auto_reset) K.reset_list[atomic_add_i32(&K.cnt[CNT_RESET], 1)] = (int32_t)env;
    }
}

// env of slot `loc` of a per-env kernel: the chunk's range, or the refill pass's list (returns -1 past the end)
MRP_HD int64_t slot_env(const SimConst& K, int64_t loc) {
    if (!K.idx_list) return loc < K.nloc ? K.env0 + loc : -1;
    const int64_t n = *K.idx_count < K.nloc ? *K.idx_count : K.nloc;
    return loc < n ? (int64_t)K.idx_list[loc] : -1;
}

// phase 0 (lane per env): broadphase half of Collide — classify every contact, queue the ones that need SAT
// returns the contacts of this env that need SAT + clipping
MRP_HD CMask broad_lane(const SimConst& K, float* sm, const float* ct, int64_t env) {
    Env e(K, sm, ct, env, nullptr, 10);
    e.load();
    return e.broad_phase();
}
// queue entries of one env, `base` = the first of its cm_count(need) reserved slots
MRP_HD void push_narrow(const SimConst& K, int64_t env, CMask need, int base) {
    while (cm_any(need)) {
        const int k = cm_first(need);
        cm_clr(need, k);
        K.narrow_list[base++] = (uint32_t)env * (uint32_t)kMaxC + (uint32_t)k;
    }
}

// phase 1 (lane per env): control + Collide events + island order + constraint setup -> solver task
// returns the number of solver constraints of the env (0: no island task was queued for it)
// staged = the warp's action rows were copied into its wall slots (k_pre, coalesced): rows are read from there, walls set afterwards
MRP_HD int pre_lane(const SimConst& K, float* sm, const float* ct, int64_t env, uint32_t* m12, const float* staged = nullptr, unsigned vmask = 0u) {
    Env e(K, sm, ct, env, nullptr, 8);
    float a[3 * MRP_MAX_AGENTS];
    if (K.hidden) {
        // the hidden step of reset(): action_space.sample() from the reset stream of the episode that just spawned (mrp00:411)
        const uint32_t episode = e.g(W_EPISODE);
        for (int i = 0; i < K.act_dim; ++i)
            a[i] = (float)(-1.0 + 2.0 * uniform53(K.seed, kStreamResetAction, e.gid, episode, (uint32_t)i));
        e.load();
    } else if (staged) {
        for (int i = 0; i < K.act_dim; ++i) a[i] = staged[i];
        e.load(false);
    } else {
        // the action row is requested first so that it is in flight while the state words are loaded
        const float* arow = K.act + env * K.act_dim;
        for (int i = 0; i < K.act_dim; ++i) a[i] = arow[i];
        e.load();
    }
    const int T = e.pre_phase(a, staged ? vmask : 0u);
    m12[0] += e.stat_m1; m12[1] += e.stat_m2;
    return T;
}

// phase 2a (lane per task): 180 velocity sweeps (early exit), StoreImpulses, position integration
struct VelTask {
    Sim::VelReg st;
    int T;
    uint32_t ops;
    uint32_t bodies;  // dynamic bodies of this island
};
MRP_HD uint32_t task_body_mask(const SimConst& K, Sim& s, int T) {
    uint32_t mask = 0;
    for (int t = 0; t < T; ++t) {
        const uint32_t vm = s.vmeta(t);
        mask |= (1u << (vm & 15)) | (1u << ((vm >> 4) & 15));
    }
    return mask & ((1u << K.nb) - 1u);
}
MRP_HD void vel_task_begin(const SimConst& K, Sim& s, VelTask& vt, int task) {
    const int64_t env = K.task_env[task];
    vt.T = K.task_T[task];
    s.set_env(env);
    s.vcp = K.pool + (size_t)K.task_off[task] * VC_WORDS;
    vt.bodies = task_body_mask(K, s, vt.T);
    for (int b = 0; b < K.nb; ++b)
        if ((vt.bodies >> b) & 1)
            for (int f = 0; f < 6; ++f) s.B(b, f) = s.gf(K.w_body + kBodyWords * b + f);
    for (int b = K.nb; b < K.nb + 4; ++b) { s.B(b, 3) = 0.0f; s.B(b, 4) = 0.0f; s.B(b, 5) = 0.0f; }
    vt.ops = 0;
}
MRP_HD void vel_task_end(const SimConst& K, Sim& s, VelTask& vt) {
#if defined(__CUDA_ARCH__)
    atomicAdd(&s.g(W_HINT), vt.ops);
#else
    s.g(W_HINT) += vt.ops;
#endif
    s.store_impulses(vt.T);
    for (int b = 0; b < K.nb; ++b) {
        if (!((vt.bodies >> b) & 1)) continue;
        s.integrate_position(b, K.h);
        for (int f = 0; f < 6; ++f) s.gsf(K.w_body + kBodyWords * b + f, s.B(b, f));
    }
}
// phase 2b (lane per task): up to 60 position sweeps
struct PosTask {
    Sim::PosState st;
    int T;
    uint32_t bodies;
};
MRP_HD void pos_task_begin(const SimConst& K, Sim& s, PosTask& pt, int task) {
    const int64_t env = K.task_env[task];
    pt.T = K.task_T[task];
    s.set_env(env);
    s.vcp = K.pool + (size_t)K.task_off[task] * VC_WORDS;
    pt.bodies = task_body_mask(K, s, pt.T);
    const float nan = s.__uint_as_float_(0x7fc00000u);
    for (int b = 0; b < K.nb; ++b) {
        if (!((pt.bodies >> b) & 1)) continue;
        for (int f = 0; f < 3; ++f) s.B(b, f) = s.gf(K.w_body + kBodyWords * b + f);
        s.set_rot_cache(b, Rot{0.0f, 1.0f}, nan);  // no rotation known for the freshly integrated angle
    }
    for (int k = 0; k < 4; ++k) {
        s.B(K.nb + k, 0) = K.ctab[CT_WALLPOS + 2 * k];
        s.B(K.nb + k, 1) = K.ctab[CT_WALLPOS + 2 * k + 1];
        s.B(K.nb + k, 2) = 0.0f;
    }
    s.pos_begin(pt.st);
}
MRP_HD void pos_task_end(const SimConst& K, Sim& s, PosTask& pt) {
    for (int b = 0; b < K.nb; ++b)
        if ((pt.bodies >> b) & 1)
            for (int f = 0; f < 3; ++f) s.gsf(K.w_body + kBodyWords * b + f, s.B(b, f));
}

// Big islands (class 3: more than two contacts; 1.7 % of the islands of a Heavy-v0 rollout, but the longest dependent
// chains of the step: up to T x 180 velocity trips and T x 60 position trips) are solved velocity AND position in one
// go by k_solve_big, a small kernel that runs beside the bulk solver kernels.  The island's constraint records are
// copied into the lane's block of shared memory first (T <= kBigT), so a trip costs shared-memory latency instead of an
// L2 round trip for records the bulk kernels keep evicting from L1, and the bodies never leave shared memory between
// the two solves.  Arithmetic and order are those of vel_task_* / pos_task_* (same Sim member functions).
constexpr int kBigT = 6;                                // contact records kept in shared memory per lane
constexpr int kBigLanes = 64;                           // lanes per CTA of k_solve_big
constexpr int kBigRecWords = kBigT * VC_WORDS + 1;      // odd: the lanes' blocks start in different banks
MRP_HD void big_task_lane(const SimConst& K, Sim& s, int task, float* rec, uint32_t& flops) {
    VelTask vt;
    vel_task_begin(K, s, vt, task);
    if (vt.T <= kBigT) {
        const float* src = s.vcp;
        for (int i = 0; i < vt.T * VC_WORDS; ++i) rec[i] = src[i];
        s.vcp = rec;
    }
    s.vr_begin(vt.st, vt.T, nullptr);
    for (;;) {
        vt.ops += (uint32_t)vt.st.vpc + 1u;
        flops += vt.st.vpc == 2 ? 160u : 81u;
        if (s.vr_trip_contact(vt.st, 180, nullptr)) break;
    }
#if defined(__CUDA_ARCH__)
    atomicAdd(&s.g(W_HINT), vt.ops);
#else
    s.g(W_HINT) += vt.ops;
#endif
    s.store_impulses(vt.T);
    const float nan = s.__uint_as_float_(0x7fc00000u);
    for (int b = 0; b < K.nb; ++b) {
        if (!((vt.bodies >> b) & 1)) continue;
        s.integrate_position(b, K.h);
        s.set_rot_cache(b, Rot{0.0f, 1.0f}, nan);  // no rotation known for the freshly integrated angle
    }
    for (int k = 0; k < 4; ++k) {
        s.B(K.nb + k, 0) = K.ctab[CT_WALLPOS + 2 * k];
        s.B(K.nb + k, 1) = K.ctab[CT_WALLPOS + 2 * k + 1];
        s.B(K.nb + k, 2) = 0.0f;
    }
    Sim::PosState ps;
    s.pos_begin(ps);
    while (!s.pos_trip<true>(ps, vt.T, 60, -1, -1)) {}
    for (int b = 0; b < K.nb; ++b)
        if ((vt.bodies >> b) & 1)
            for (int f = 0; f < 6; ++f) s.gsf(K.w_body + kBodyWords * b + f, s.B(b, f));
}

// phase 3 (lane per env): transforms, broadphase, TOI, obs / reward / done, TimeLimit.  With allow_events ==
// false an env whose TOI scan finds an event is queued for the event pass and left untouched.
#ifdef MRP_TAILPROBE
// per-env breakdown of the event pass: {total, inside time_of_impact, inside toi_event} cycles, TOI calls, events
__device__ unsigned long long g_tp_evrec[65536][5];
__device__ unsigned int g_tp_evn;
#endif
MRP_HD void post_lane(const SimConst& K, float* sm, const float* ct, int64_t env, bool allow_events, float* vc_local,
                      bool free_group = false, unsigned entry = 0u) {
    Env e(K, sm, ct, env, vc_local, allow_events ? kDynFields : 11);
#if defined(MRP_TAILPROBE) && defined(__CUDA_ARCH__)
    const long long tp_c0 = clock64();
#endif
    e.load();
    double r;
    bool d;
    const bool fin = e.post_phase(K.obs + env * K.obs_dim, &r, &d, allow_events, entry);
#if defined(MRP_TAILPROBE) && defined(__CUDA_ARCH__)
    if (allow_events) {
        const unsigned int i = atomicAdd(&g_tp_evn, 1u);
        if (i < 65536u) {
            g_tp_evrec[i][0] = (unsigned long long)(clock64() - tp_c0); g_tp_evrec[i][1] = (unsigned long long)e.tp_toi_clk;
            g_tp_evrec[i][2] = (unsigned long long)e.tp_evt_clk; g_tp_evrec[i][3] = e.stat_toi; g_tp_evrec[i][4] = (unsigned long long)e.tp_evt_n;
        }
    }
#endif
    if (!fin) {
        // deferred to the event pass; the task-free group has its own queue (end of toi_list, downwards) so that its
        // events can be processed beside the solver kernels as well
        if (free_group) K.toi_list[K.nloc - 1 - atomic_add_i32(&K.cnt[CNT_TOI_F], 1)] = (int32_t)env;
        else K.toi_list[atomic_add_i32(&K.cnt[CNT_TOI], 1)] = (int32_t)env;
        return;
    }
    if (K.hidden) {   // reset_env: the hidden step's reward and done flag are dropped, the state is the episode's first
        if (!e.stored) e.store();
        K.spare_ok[env] = 1;
        return;
    }
    finish_step(K, e, env, d, r);
}

// fused single-lane step (used by the host emulation's reference path and kept for debugging)
// Constraint records of the fused per-env paths (fused step, TOI event pass, reset): a lane-local array in the default
// build; the wide build (192 x 38 words) borrows the env's own slice of the task pool.  The solver records of a step are
// bump-allocated from the start of the same pool (k_pre), so these paths must not run concurrently with k_solve_vel /
// k_solve_pos of the same chunk: launch_step() orders the wide build's event passes after the solver kernels.
#ifdef MRP_WIDE
#define MRP_VC_SCRATCH(K, env) float* vc_local = (K).pool + (size_t)((env) - (K).env0) * (K).maxc * VC_WORDS
#else
#define MRP_VC_SCRATCH(K, env) float vc_local[kMaxC * VC_WORDS]
#endif

MRP_HD void step_lane(const SimConst& K, float* sm, const float* ct, int64_t env) {
    MRP_VC_SCRATCH(K, env);
    Env e(K, sm, ct, env, vc_local);
    e.load();
    float a[3 * MRP_MAX_AGENTS];
    const float* arow = K.act + env * K.act_dim;
    for (int i = 0; i < K.act_dim; ++i) a[i] = arow[i];
    double r;
    bool d = e.env_step(a, K.obs + env * K.obs_dim, &r, false);
    finish_step(K, e, env, d, r);
}

MRP_HD void reset_lane(const SimConst& K, float* sm, const float* ct, int64_t env) {
    MRP_VC_SCRATCH(K, env);
    Env e(K, sm, ct, env, vc_local);
    e.reset_env(K.obs + env * K.obs_dim);
    if (e.overflow) stat_add(K.stats, MRP_STAT_OVERFLOW, 1.0);
}

// first half of reset_env for the refill pass: episode counters, spawn, contacts of the fresh fixtures; the hidden step follows
// as an ordinary pipeline pass over the same list (K2: S = S2, hidden)
MRP_HD void spawn_spare_lane(const SimConst& K, const SimConst& K2, float* sm, const float* ct, int64_t env) {
    Env e(K2, sm, ct, env, nullptr);
    const uint32_t episode = env_words(K, env)[W_EPISODE << kTileShift] + 1u;
    e.g(W_EPISODE) = episode;
    e.g(W_ELAPSED) = 0;
    e.g(W_EPLEN) = 0;
    e.gsd(W_EPRET, 0.0);
    e.spawn(episode);
    e.find_new_contacts(0xffffffffu);
    e.store();
    // the body origins as spawned, for the collide half of the hidden step (Sim::load(), narrow_item)
    for (int b = 0; b < K.nb; ++b) {
        e.gsf(K.w_body + kBodyWords * b + 8, e.BX(b, 8));
        e.gsf(K.w_body + kBodyWords * b + 9, e.BX(b, 9));
    }
    if (e.overflow) stat_add(K.stats, MRP_STAT_OVERFLOW, 1.0);
}
// auto-reset of a finished env from its spare: every state word and the observation row.  Returns false when there is no valid spare.
MRP_HD bool reset_from_spare(const SimConst& K, int64_t env) {
    if (!K.S2 || !K.spare_ok[env]) return false;
    const uint32_t* src = K.S2 + (env >> kTileShift) * ((int64_t)K.w_total << kTileShift) + (env & (kTile - 1));
    uint32_t* dst = env_words(K, env);
    const uint32_t hint = dst[W_HINT << kTileShift];
    for (int w = 0; w < K.w_total; ++w) dst[w << kTileShift] = src[w << kTileShift];
    dst[W_HINT << kTileShift] = hint;   // ordering hint of the solver queues: not part of the episode
    const float* so = K.obs2 + env * K.obs_dim;
    float* o = K.obs + env * K.obs_dim;
    for (int i = 0; i < K.obs_dim; ++i) o[i] = so[i];
    K.spare_ok[env] = 0;
    return true;
}

MRP_HD void sample_actions_lane(const SimConst& K, float* dst, uint64_t step_index, int64_t env) {
    uint64_t gid = K.env_id_base + (uint64_t)env;
    for (int k = 0; k < K.act_dim; ++k)
        dst[env * K.act_dim + k] = (float)(-1.0 + 2.0 * uniform53(K.seed, kStreamAction, gid, (uint32_t)step_index, (uint32_t)k));
}

// Alternative observation head: the normalised observation of the reference's experimental MultiRobotPuzzle-v3
// (gym_puzzles/envs/core.py:289-350, _get_norm_pose / _get_obs), computed from the state of a v0-family env:
// per robot [bx - ax, by - ay, rot mod 2pi, contact], block [gx - bx, gy - by, grot - brot], 8 block vertices; positions
// (x - w/2) / (w/2), (y - h/2) / (w/2) with w, h the arena size in metres (core.py divides both by the WIDTH scale).
MRP_HD void obs_v3_lane(const SimConst& K, float* out, int64_t env) {
    const uint32_t* G = env_words(K, env);
    auto gfl = [&](int w) { union { uint32_t u; float f; } c; c.u = G[w << kTileShift]; return c.f; };
    auto gdb = [&](int w) { union { uint64_t u; double d; } c; c.u = (uint64_t)G[w << kTileShift] | ((uint64_t)G[(w + 1) << kTileShift] << 32); return c.d; };
    const double ws = K.W / 2, hs = K.H / 2;
    float* o = out + env * (4 * K.n + 19);
    const int wb = K.w_body;
    const double bx = ((double)gfl(wb) - ws) / ws, by = ((double)gfl(wb + 1) - hs) / ws, brot = py_mod((double)gfl(wb + 2), kTwoPiD);
    const uint32_t goalc = G[W_GOALC << kTileShift];
    for (int i = 0; i < K.n; ++i) {
        const int w = wb + kBodyWords * (1 + i);
        const double ax = ((double)gfl(w) - ws) / ws, ay = ((double)gfl(w + 1) - hs) / ws;
        *o++ = (float)(bx - ax);
        *o++ = (float)(by - ay);
        *o++ = (float)py_mod((double)gfl(w + 2), kTwoPiD);
        *o++ = ((goalc >> i) & 1) ? 1.0f : 0.0f;
    }
    const double sw = K.W * K.SCALE, sh = K.H * K.SCALE;            // screen size in pixels (640 x 480)
    const double gx = (gdb(W_GOAL) - sw / 2) / (sw / 2), gy = (gdb(W_GOAL + 2) - sh / 2) / (sw / 2);
    *o++ = (float)(gx - bx);
    *o++ = (float)(gy - by);
    *o++ = (float)(0.0 - brot);
    Xf xf;
    xf.q.s = gfl(wb + 6); xf.q.c = gfl(wb + 7);
    const V2 r = rmul(xf.q, mk(K.blk_lcx, K.blk_lcy));
    xf.p = mk(gfl(wb) - r.x, gfl(wb + 1) - r.y);
    for (int k = 0; k < 8; ++k) {
        const V2 p = xmul(xf, mk(K.blkv[k][0], K.blkv[k][1]));
        *o++ = (float)(((double)p.x - ws) / ws);
        *o++ = (float)(((double)p.y - hs) / ws);
    }
}

MRP_HD void fix_rot_lane(const SimConst& K, int64_t env) {  // q = Rot(a) after a state upload
    for (int b = 0; b < K.nb; ++b) {
        uint32_t* G = env_words(K, env);
        union { uint32_t u; float f; } c;
        c.u = G[(K.w_body + kBodyWords * b + 2) << kTileShift];
        Rot q = rot_set(c.f);
        c.f = q.s; G[(K.w_body + kBodyWords * b + 6) << kTileShift] = c.u;
        c.f = q.c; G[(K.w_body + kBodyWords * b + 7) << kTileShift] = c.u;
    }
}

#ifndef MRP_HOST_EMU
__device__ __forceinline__ const float* load_ctab(const SimConst& K, float* smem) {
    for (int i = threadIdx.x; i < CT_WORDS; i += blockDim.x) smem[i] = K.ctab[i];
    __syncthreads();
    return smem;
}

// -DMRP_TAILPROBE (profiling builds only, profiles/tailprobe.py): when does each warp of the tail-dominated kernels finish,
// and which single task ran longest?  Times are globaltimer nanoseconds.
#ifdef MRP_TAILPROBE
constexpr int kTpKernels = 8, kTpWarps = 8192;   // ids: 0 k_solve_vel, 1 k_solve_pos, 2 k_post_events, 3 k_post_events (free), 4..7 vel classes 0..3
__device__ unsigned long long g_tp_start[kTpKernels];
__device__ unsigned long long g_tp_end[kTpKernels][kTpWarps];
__device__ unsigned long long g_tp_maxtask[kTpKernels];
__device__ __forceinline__ unsigned long long tp_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void tp_begin(int id) { if ((threadIdx.x & 31) == 0) atomicMin(&g_tp_start[id], tp_now()); }
__device__ __forceinline__ void tp_end(int id) {
    const int w = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if ((threadIdx.x & 31) == 0 && w < kTpWarps) g_tp_end[id][w] = tp_now();
}
__device__ __forceinline__ void tp_task(int id, unsigned long long t0, uint32_t info) {
    atomicMax(&g_tp_maxtask[id], ((tp_now() - t0) << 24) | (unsigned long long)(info & 0xffffffu));
}
#define TP(x) x
#else
#define TP(x)
#endif

// this lane's column of the CTA's per-lane shared memory: every warp owns words_per_lane x 32 floats
__device__ __forceinline__ float* lane_sm(float* base, int words_per_lane) {
    return base + (threadIdx.x >> 5) * (words_per_lane * 32) + (threadIdx.x & 31);
}

__global__ void k_clear(int32_t* cnt) {
    if (threadIdx.x < CNT_N) cnt[threadIdx.x] = 0;
}
static_assert(CNT_N <= 32, "k_clear is launched with 32 threads");

// fused single-kernel step (MRP_FUSED_STEP=1): kept for A/B measurements against the phase pipeline
__global__ void __launch_bounds__(kBlock) k_step(const __grid_constant__ SimConst K) {
    extern __shared__ float smem[];
    const float* ct = load_ctab(K, smem);
    const int64_t loc = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (loc >= K.nloc) return;
    const int64_t env = K.env0 + loc;
    step_lane(K, lane_sm(smem + kCtPad, K.smem_words), ct, env);
}

__global__ void __launch_bounds__(kBlock) k_broad(const __grid_constant__ SimConst K) {
    extern __shared__ float smem[];
    const float* ct = load_ctab(K, smem);
    const int64_t loc = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t env = slot_env(K, loc);
    const bool valid = env >= 0;   // every lane of the warp stays for the warp-aggregated queue reservation
    CMask need = cm_none();
    if (valid) need = broad_lane(K, lane_sm(smem + kCtPad, 10 * K.nb + 4 * K.ndynfix), ct, env);
    // one atomicAdd per warp instead of one per queued contact (they were 14 % of this kernel's stall samples)
    const int n = cm_count(need), lane = threadIdx.x & 31;
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(&K.cnt[CNT_NARROW], total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (n) push_narrow(K, env, need, base + incl - n);
}

// SAT + clipping for the queued contacts: one lane per contact, grid-stride over the queue
__global__ void __launch_bounds__(kBlock) k_narrow(const __grid_constant__ SimConst K) {
    extern __shared__ float smem[];
    const int count = K.cnt[CNT_NARROW];
    if ((int64_t)blockIdx.x * kBlock >= count) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(K.stats + MRP_STAT_PAIRS, (double)count);
    const float* ct = load_ctab(K, smem);
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < count; i += (int64_t)gridDim.x * kBlock)
        narrow_item(K, ct, K.narrow_list[i]);
}

// envs [loc0, loc1) of the chunk: mrp_step_host launches it in two halves so that the first half starts as soon as
// its actions have arrived
__global__ void __launch_bounds__(kBlock) k_pre(const __grid_constant__ SimConst K, int64_t loc0, int64_t loc1) {
    extern __shared__ float smem[];
    const float* ct = load_ctab(K, smem);
    const int64_t loc = loc0 + (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t env = loc < loc1 ? slot_env(K, loc) : -1;
    const bool valid = env >= 0;   // every lane of the warp stays for the warp-aggregated list reservation
    int T = -1;
    uint32_t m12[2] = {0u, 0u};
    float* const lsm = lane_sm(smem + kCtPad, 8 * K.nb + 24);
    const float* staged = nullptr;
    if (K.stage_rows && !K.hidden) {
        // the 32 action rows of the warp's envs are one contiguous run of 32 x act_dim floats: consecutive lanes fetch
        // consecutive words (coalesced) into the warp's wall slots (24 words per lane, set only after the rows have been read),
        // then every lane picks up its own row
        const int lane = threadIdx.x & 31;
        float* blk = lsm - lane + (8 * K.nb) * 32;
        const int64_t first = loc - lane;
        const int words = (int)((loc1 - first < 32 ? loc1 - first : 32) * K.act_dim);
        const float* src = K.act + (K.env0 + first) * K.act_dim;
        for (int i = lane; i < words; i += 32) blk[i] = src[i];
        __syncwarp();
        staged = blk + lane * K.act_dim;
    }
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) T = pre_lane(K, lsm, ct, env, m12, staged, vmask);
    stat_add_warp(K.stats, MRP_STAT_M1, m12[0]);
    stat_add_warp(K.stats, MRP_STAT_M2, m12[1]);
    // envs without solver tasks are final already: k_post handles them while the solver kernels run (post_list)
    const unsigned free_m = __ballot_sync(0xffffffffu, T == 0), busy_m = __ballot_sync(0xffffffffu, T > 0);
    const int lane = threadIdx.x & 31;
    int base_f = 0, base_b = 0;
    if (lane == 0) {
        if (free_m) base_f = atomicAdd(&K.cnt[CNT_FREE], __popc(free_m));
        if (busy_m) base_b = atomicAdd(&K.cnt[CNT_BUSY], __popc(busy_m));
    }
    base_f = __shfl_sync(0xffffffffu, base_f, 0);
    base_b = __shfl_sync(0xffffffffu, base_b, 0);
    const unsigned below = (1u << lane) - 1u;
    if (T == 0) K.post_list[base_f + __popc(free_m & below)] = (int32_t)env;
    else if (T > 0) K.post_list[K.nloc - 1 - (base_b + __popc(busy_m & below))] = (int32_t)env;
}

// Persistent solver kernels.  Every lane owns one task at a time and advances it by one point operation per loop
// trip; a warp refills its idle lanes from the global task queue once kRefill of them are idle, so the 32 lanes
// stay busy although islands need anywhere between 2 and ~1000 operations.
#ifndef MRP_REFILL
#define MRP_REFILL 8
#endif
constexpr int kRefill = MRP_REFILL;
#ifndef MRP_INNER_TRIPS
#define MRP_INNER_TRIPS 8
#endif
constexpr int kInnerTrips = MRP_INNER_TRIPS;

template <int CLS>
__device__ __forceinline__ void solve_vel_class(const SimConst& K, Sim& s, uint32_t& flops) {
    const int ntasks = task_count(K, CLS);
    VelTask vt;
    Sim::VelReg st1;  // second contact (class 2 only)
    float imp3[CLS == 3 ? 4 * kMaxC : 4];   // class 3: the island's accumulated impulses while it iterates
    float* const imp = CLS == 3 ? imp3 : nullptr;
    bool busy = false, exhausted = false;
    TP(unsigned long long tp_t0 = 0;)
    for (;;) {
        const unsigned bm = __ballot_sync(0xffffffffu, busy);
        if (32 - __popc(bm) >= kRefill || bm == 0u) {
            if (!busy && !exhausted) {
                const int task = atomicAdd(&K.cnt[CNT_HEAD_V + CLS], 1);
                if (task < ntasks) {
                    TP(tp_t0 = tp_now();)
                    vel_task_begin(K, s, vt, task_slot(K, CLS, task));
                    if (CLS == 2) s.vr_begin_pair(vt.st, st1); else s.vr_begin(vt.st, vt.T, imp);
                    busy = true;
                } else exhausted = true;
            }
            if (__ballot_sync(0xffffffffu, busy) == 0u) break;
        }
        if (busy) {
            // a few trips between two refill checks: the warp-wide ballots (reconvergence points) were 17 % of this
            // kernel's stall samples when taken every trip
            bool fin = false;
#pragma unroll 1
            for (int rep = 0; rep < kInnerTrips && !fin; ++rep) {
                // flops by the cost model of SURVEY.md Appendix D: 81 per 1-point contact and sweep, 160 per 2-point contact
                if (CLS == 0) { vt.ops += 2; flops += 81u; fin = s.vr_sweep_single<1>(vt.st, 180); }
                else if (CLS == 1) { vt.ops += 3; flops += 160u; fin = s.vr_sweep_single<2>(vt.st, 180); }
                else if (CLS == 2) { vt.ops += 5; flops += (vt.st.vpc == 2 ? 160u : 81u) + (st1.vpc == 2 ? 160u : 81u); fin = s.vr_sweep_pair(vt.st, st1, 180); }
                else { vt.ops += (uint32_t)vt.st.vpc + 1u; flops += vt.st.vpc == 2 ? 160u : 81u; fin = s.vr_trip_contact(vt.st, 180, imp); }
            }
            if (fin) {
                TP(tp_task(0, tp_t0, ((uint32_t)vt.T << 16) | (uint32_t)(CLS == 2 ? vt.st.sweep : vt.st.sweep));)
                vel_task_end(K, s, vt);
                busy = false;
            }
        }
    }
    TP(tp_end(4 + CLS);)
}

__global__ void __launch_bounds__(kBlock) k_solve_vel(const __grid_constant__ SimConst K) {
    extern __shared__ float smem[];
    Sim s(K, lane_sm(smem, 6 * (K.nb + 4)), K.ctab, 0, nullptr, 6);
    // every warp serves one class at a time (uniform instruction stream); multi-contact islands first, they run longest
    uint32_t flops = 0u;
    TP(tp_begin(0); tp_begin(4); tp_begin(5); tp_begin(6); tp_begin(7);)
    if (!K.big_split) solve_vel_class<3>(K, s, flops);
    solve_vel_class<2>(K, s, flops);
    solve_vel_class<1>(K, s, flops);
    solve_vel_class<0>(K, s, flops);
    stat_add_warp(K.stats, MRP_STAT_VEL_FLOPS, flops);
    TP(tp_end(0);)
}

// lane per big island, one task at a time; the lanes of a warp walk through velocity solve, then position solve together
__global__ void __launch_bounds__(kBigLanes) k_solve_big(const __grid_constant__ SimConst K) {
    extern __shared__ float smem[];
    const int bw = 9 * K.nb + 24;
    Sim s(K, lane_sm(smem, bw), K.ctab, 0, nullptr, 9);
    float* rec = smem + (kBigLanes >> 5) * bw * 32 + threadIdx.x * kBigRecWords;
    const int ntasks = task_count(K, 3);
    uint32_t flops = 0u;
    TP(tp_begin(7);)
    for (;;) {
        const int task = atomicAdd(&K.cnt[CNT_HEAD_V + 3], 1);
        if (task >= ntasks) break;
        TP(const unsigned long long tp_t0 = tp_now();)
        big_task_lane(K, s, task_slot(K, 3, task), rec, flops);
        TP(tp_task(7, tp_t0, (uint32_t)K.task_T[task_slot(K, 3, task)] << 16);)
    }
    TP(tp_end(7);)
    stat_add_warp(K.stats, MRP_STAT_VEL_FLOPS, flops);
    stat_add_warp(K.stats, MRP_STAT_POS_POINTS, s.stat_pos_pts);
}

__global__ void __launch_bounds__(kBlock) k_solve_pos(const __grid_constant__ SimConst K) {
    extern __shared__ float smem[];
    Sim s(K, lane_sm(smem, 9 * K.nb + 24), K.ctab, 0, nullptr, 9);
    const int ntasks = task_count_all(K);
    PosTask pt;
    bool busy = false, exhausted = false;
    TP(unsigned long long tp_t0 = 0; tp_begin(1);)
    for (;;) {
        const unsigned bm = __ballot_sync(0xffffffffu, busy);
        if (32 - __popc(bm) >= kRefill || bm == 0u) {
            if (!busy && !exhausted) {
                const int task = atomicAdd(&K.cnt[CNT_HEAD_P], 1);
                if (task < ntasks) { TP(tp_t0 = tp_now();) pos_task_begin(K, s, pt, task_slot_any(K, task)); busy = true; }
                else exhausted = true;
            }
            if (__ballot_sync(0xffffffffu, busy) == 0u) break;
        }
        if (busy) {
            bool fin = false;
#pragma unroll 1
            for (int rep = 0; rep < kInnerTrips && !fin; ++rep) fin = s.pos_trip<true>(pt.st, pt.T, 60, -1, -1);
            if (fin) {
                TP(tp_task(1, tp_t0, ((uint32_t)pt.T << 16) | (uint32_t)pt.st.sweep);)
                pos_task_end(K, s, pt);
                busy = false;
            }
        }
    }
    stat_add_warp(K.stats, MRP_STAT_POS_POINTS, s.stat_pos_pts);
    TP(tp_end(1);)
}

// which = 2: envs [env0, env0 + nloc) in order.  which = 0 / 1: the envs k_pre listed as free of / owning solver tasks
// (post_list): the free ones do not depend on the solver kernels and run beside them on a second stream.
__global__ void __launch_bounds__(kBlock, 4) k_post(const __grid_constant__ SimConst K, int which) {
    extern __shared__ float smem[];
    const int64_t loc = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    int64_t count = which == 2 ? (int64_t)K.nloc : (int64_t)K.cnt[which == 0 ? CNT_FREE : CNT_BUSY];
    if (which == 2 && K.idx_list && *K.idx_count < count) count = *K.idx_count;
    if ((int64_t)blockIdx.x * kBlock >= count) return;
    const float* ct = load_ctab(K, smem);
    const unsigned entry = K.stage_rows ? __ballot_sync(0xffffffffu, loc < count) : 0u;   // lanes of this warp that own an env
    if (loc >= count) return;
    const int64_t env = which == 2 ? slot_env(K, loc) : (int64_t)K.post_list[which == 0 ? loc : K.nloc - 1 - loc];
    post_lane(K, lane_sm(smem + kCtPad, 11 * K.nb + 4 * K.ndynfix), ct, env, false, nullptr, which == 0, entry);
}

// rare paths, grid-stride over their queues: envs with a TOI event this step; envs to auto-reset
__global__ void __launch_bounds__(kBlock) k_post_events(const __grid_constant__ SimConst K, int free_group) {
    extern __shared__ float smem[];
    const int count = K.cnt[free_group ? CNT_TOI_F : CNT_TOI];
    if ((int64_t)blockIdx.x * kBlock >= count) return;
    const float* ct = load_ctab(K, smem);
    TP(tp_begin(2 + (free_group ? 1 : 0));)
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < count; i += (int64_t)gridDim.x * kBlock) {
        const int64_t env = K.toi_list[free_group ? K.nloc - 1 - i : i];
        MRP_VC_SCRATCH(K, env);
        TP(const unsigned long long tp_t0 = tp_now();)
        post_lane(K, lane_sm(smem + kCtPad, K.smem_words), ct, env, true, vc_local);
        TP(tp_task(2 + (free_group ? 1 : 0), tp_t0, 0u);)
    }
    TP(tp_end(2 + (free_group ? 1 : 0));)
}

// `lanes` lanes per warp take an env (a power of two), the others idle: a respawned env runs a whole fused step — its own
// path through collide, the solver sweeps and the TOI loop — so the lanes of a warp execute their streams one after the other,
// and the queue is short (envs whose episode ended this step)
__global__ void __launch_bounds__(kBlock) k_reset_list(const __grid_constant__ SimConst K, int lanes) {
    extern __shared__ float smem[];
    const int count = K.cnt[CNT_RESET];
    if ((int64_t)blockIdx.x * (kBlock / 32) * lanes >= count) return;
    const float* ct = load_ctab(K, smem);
    const int lane = threadIdx.x & 31, step = 32 / lanes;
    if (lane % step) return;
    const int64_t warp = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5, stride = (int64_t)gridDim.x * (kBlock / 32) * lanes;
    for (int64_t i = warp * lanes + lane / step; i < count; i += stride) {
        const int64_t env = K.reset_list[i];
        if (!reset_from_spare(K, env)) reset_lane(K, lane_sm(smem + kCtPad, K.smem_words), ct, env);
    }
}

// spare episodes: list the envs whose spare is missing (one warp-aggregated atomic per warp) ...
__global__ void k_refill_collect(const __grid_constant__ SimConst K) {
    // 16 flags per thread (one 16-byte load; nearly all are 1 in steady state)
    const int64_t first = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    uint32_t miss = 0;
    if (first < K.N) {
        const uint4 f = *reinterpret_cast<const uint4*>(K.spare_ok + first);
        const uint32_t w[4] = {f.x, f.y, f.z, f.w};
        for (int i = 0; i < 16; ++i)
            if (first + i < K.N && !((w[i >> 2] >> (8 * (i & 3))) & 0xffu)) miss |= 1u << i;
    }
    const int n = __popc(miss);
    if (!__ballot_sync(0xffffffffu, n > 0)) return;
    const int lane = threadIdx.x & 31;
    int incl = n;
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    int base = 0;
    if (lane == 31) base = atomicAdd(K.refill_cnt, incl);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - n;
    for (; miss; miss &= miss - 1u, ++base)
        if (base < K.refill_cap) K.refill_list[base] = (int32_t)(first + __ffs((int)miss) - 1);   // the rest gets its spare in a later step
}
// ... and start their next episode in the spare buffers (spawn_spare_lane); the hidden step then runs as a pipeline pass over the
// same list on the same low-priority stream (launch_refill)
__global__ void __launch_bounds__(kBlock) k_spawn_list(const __grid_constant__ SimConst K, const __grid_constant__ SimConst K2) {
    extern __shared__ float smem[];
    const int64_t loc = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t n = *K2.idx_count < K2.nloc ? *K2.idx_count : K2.nloc;
    if ((int64_t)blockIdx.x * kBlock >= n) return;
    const float* ct = load_ctab(K, smem);
    if (loc >= n) return;
    spawn_spare_lane(K, K2, lane_sm(smem + kCtPad, K.smem_words), ct, K2.idx_list[loc]);
}

// mrp_step_host with a pinned, device-visible obs buffer: the rows of a chunk leave by cudaMemcpyAsync as soon as its k_post
// has finished, i.e. before the chunk's TOI-event and auto-reset passes (a serial tail of ~0.4 ms that touches ~3 % of the
// envs); the rows those two passes rewrite are then stored straight into the host buffer by this kernel (zero-copy stores
// over PCIe: 47 GB/s for scattered rows, profiles/micro/zerocopy_rows.cu), ordered after the bulk copy of the same range.
// which 0: the chunk's TOI-event queue, 1: its auto-reset queue.  V = float4 when the row length allows it, else float.
template <typename V>
__global__ void __launch_bounds__(256) k_out_rows(const __grid_constant__ SimConst K, int which, V* __restrict__ obs_host) {
    const int64_t count = K.cnt[which == 0 ? CNT_TOI : CNT_RESET];
    const int32_t* list = which == 0 ? K.toi_list : K.reset_list;
    const int q = K.obs_dim / (int)(sizeof(V) / sizeof(float));   // elements per row
    const int64_t total = count * q;
    const V* src = reinterpret_cast<const V*>(K.obs);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / q;
        const int64_t at = (int64_t)list[r] * q + (i - r * q);
        obs_host[at] = src[at];
    }
}

__global__ void __launch_bounds__(kBlock) k_reset_mask(const __grid_constant__ SimConst K) {
    extern __shared__ float smem[];
    const float* ct = load_ctab(K, smem);
    int64_t env = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (env >= K.N) return;
    if (K.reset_mask && !K.reset_mask[env]) return;
    reset_lane(K, lane_sm(smem + kCtPad, K.smem_words), ct, env);
    if (K.spare_ok) K.spare_ok[env] = 0;   // the spare was the episode this call just started
}

__global__ void k_sample_actions(const __grid_constant__ SimConst K, float* dst, uint64_t step_index) {
    int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env < K.N) sample_actions_lane(K, dst, step_index, env);
}

__global__ void k_obs_v3(const __grid_constant__ SimConst K, float* out) {
    int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env < K.N) obs_v3_lane(K, out, env);
}

__global__ void k_fix_rot(const __grid_constant__ SimConst K, int64_t begin, int64_t count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) fix_rot_lane(K, begin + i);
}
#endif

}  // namespace

// ------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------
constexpr int kMaxChunks = 8;
constexpr int kMaxWaves = 4;   // mrp_step_host: the env range can run as up to four front-half waves (see there)

struct mrp_handle {
#ifndef MRP_WIDE
    struct mrp_handle_wide* wide;   // set: this handle only fronts a wide-capacity handle (every call forwards to it)
#endif
    SimConst K;
    mrp_layout L;
    int device;
    float* ctab_dev;
    float ctab_host[CT_WORDS];   // host copy of the constant table (fixture -> body map for mrp_set_state)
    float* act_dev;
    uint8_t* mask_dev;  // mrp_reset_host: staging buffer of the host mask (allocated on first use)
    int fused;        // MRP_FUSED_STEP=1: single fused kernel per step (debug / A-B comparison)
    size_t smem_vel, smem_pos, smem_broad, smem_pre, smem_post, smem_big;
    int big_ctas;     // CTAs of k_solve_big per SM (MRP_BIG_CTAS)
    int big_split;    // islands with more than two contacts go to k_solve_big on a side stream (MRP_BIG, default: from 32768 envs)
    int solver_ctas;  // persistent solver CTAs per SM
    int reset_lanes;  // MRP_RESET_LANES: lanes per warp that take an env in the auto-reset / spare-episode passes
    // queues / pool / counters of the refill pass (a second, smaller set: at most refill_cap envs per step)
    int32_t *r_cnt, *r_task_env, *r_task_T, *r_task_off, *r_toi, *r_post;
    uint32_t* r_narrow;
    float* r_pool;
    double* r_stats;
    int use_spares;   // MRP_SPARES (default: from 131,072 envs with auto-reset, capacity-32 build): next episodes computed ahead of time
    int num_sms;      // multiprocessors of the handle's device (148 on B200); persistent / queue grids are sized from it
    int nchunks;       // mrp_step: the env range runs as nchunks independent pipelines on separate streams
    int nchunks_host;  // mrp_step_host: same, with each chunk's H2D / D2H copies on its stream
#ifndef MRP_HOST_EMU
    cudaStream_t cstream[kMaxChunks];
    cudaEvent_t cfork, cact, cact0, cpre, cfree, cbig, cjoin[kMaxChunks], cpost[kMaxChunks], cd2h[kMaxChunks], cdone[kMaxChunks];
    cudaEvent_t wpre[kMaxWaves], wbig[kMaxWaves], wjoin[kMaxWaves];
    int host_waves;            // MRP_HOST_WAVES: front-half waves of mrp_step_host (default: measured best per batch size)
    int wave_bound[kMaxWaves + 1];   // wave w covers the envs of back chunks [wave_bound[w], wave_bound[w + 1]) (of nchunks_host)
    cudaStream_t copy_stream;  // mrp_step_host: bulk obs copies of the chunks (early-copy path)
    cudaStream_t rstream;      // spare episodes: the refill pass beside the step's kernels (highest priority, small grids)
    int refill_pending;        // a refill pass is in flight on rstream (joined before the auto-resets of the step)
    int32_t* refill_seen;      // pinned: length of the refill list of an earlier step (sizes the pass's grids without a sync)
    cudaEvent_t crf0, crf1;
    int host_early_copy;       // MRP_HOST_EARLY_COPY (default 1): copy a chunk's rows before its event / reset passes
    cudaEvent_t tr[32];        // MRP_TRACE=1: timeline of one mrp_step_host call (created on first use)
    int tr_init;
    const void* zc_host;       // last obs_host pointer checked for device visibility ...
    float* zc_dev;             // ... and its device alias (nullptr: not pinned / not mapped -> rows leave after the passes)
    int overlap_post;  // mrp_step: k_post of envs without solver tasks runs beside the solver kernels
    // Small batches are bound by the launch chain (eleven launches and up to five copies per step, each a few microseconds of
    // host time for kernels that run as long): a step whose launches all sit on one stream is captured once into a CUDA graph
    // and replayed with a single cudaGraphLaunch.  The kernel parameters are baked into the graph, so it is re-captured when
    // the handle's constants (reward parameters, optional buffers) or the caller's pointers change.
    int use_graph;               // MRP_GRAPH (default: batches below 32,768 envs, where no side stream is used)
    cudaGraphExec_t gx_step, gx_host;
    SimConst gk_step, gk_host;   // handle constants the graphs were captured with
    const void* gp_host[5];      // host pointers of the captured mrp_step_host
    const float* gp_act;         // action pointer of the captured mrp_step
    int64_t gl_step, gl_host;    // launches inside one replay
#endif
    int64_t launches;
    int64_t steps_done;  // mrp_step / mrp_step_host calls since the statistics were last reset (env_steps = steps_done * N)
    size_t smem_bytes;
    // optional device timing of the step kernel alone (bench.py roofline): ring of event pairs
    int timing;
    int ev_n;          // pairs recorded and not yet accumulated
    double k_ms;       // accumulated k_step milliseconds
    int64_t k_count;   // accumulated k_step launches
#ifndef MRP_HOST_EMU
    cudaEvent_t ev0[64], ev1[64];
    cudaEvent_t evk[64][4];   // boundaries between the five phase kernels of one step
    double phase_ms[5];       // accumulated per-kernel milliseconds: pre, solve_vel, solve_pos, post, post_events
#endif
#ifdef MRP_HOST_EMU
    float* emu_sm;
#endif
};

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

#ifdef MRP_HOST_EMU
#define DEV_ALLOC(ptr, bytes) ((*(void**)&(ptr) = calloc(1, (bytes))) ? 0 : -1)
#define DEV_ALLOC_RAW(ptr, bytes) ((*(void**)&(ptr) = malloc((bytes))) ? 0 : -1)
#define DEV_FREE(ptr) free(ptr)
#define H2D(dst, src, bytes) (memcpy((dst), (src), (bytes)), 0)
#define D2H(dst, src, bytes) (memcpy((dst), (src), (bytes)), 0)
#define DEV_ZERO(ptr, bytes) (memset((ptr), 0, (bytes)), 0)
static const char* dev_err() { return "host allocation failed"; }
#else
#define DEV_ALLOC(ptr, bytes) (cudaMalloc((void**)&(ptr), (bytes)) == cudaSuccess ? (cudaMemset((ptr), 0, (bytes)), 0) : -1)
#define DEV_ALLOC_RAW(ptr, bytes) (cudaMalloc((void**)&(ptr), (bytes)) == cudaSuccess ? 0 : -1)
#define DEV_FREE(ptr) cudaFree(ptr)
#define H2D(dst, src, bytes) (cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1)
#define D2H(dst, src, bytes) (cudaMemcpy((dst), (src), (bytes), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1)
#define DEV_ZERO(ptr, bytes) (cudaMemset((ptr), 0, (bytes)) == cudaSuccess ? 0 : -1)
static const char* dev_err() { return cudaGetErrorString(cudaGetLastError()); }
static int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return -10;
    }
    return 0;
}
#endif

#ifndef MRP_WIDE
// entry points of the wide-capacity compilation of this file (same signatures, suffix _wide)
extern "C" {
const char* mrp_last_error_wide(void);
int mrp_destroy_wide(mrp_handle_wide*);
int mrp_create_wide(const mrp_config*, mrp_handle_wide**);
int mrp_get_layout_wide(mrp_handle_wide*, mrp_layout*);
int mrp_get_buffers_wide(mrp_handle_wide*, mrp_buffers*);
int mrp_reset_wide(mrp_handle_wide*, const uint8_t*, void*);
int mrp_set_timing_wide(mrp_handle_wide*, int32_t);
int mrp_get_timing_wide(mrp_handle_wide*, double*, int64_t*, int32_t);
int mrp_get_phase_timing_wide(mrp_handle_wide*, double*, int32_t);
int mrp_step_wide(mrp_handle_wide*, const float*, void*);
int mrp_step_host_wide(mrp_handle_wide*, const float*, float*, float*, uint8_t*, uint8_t*);
int mrp_reset_host_wide(mrp_handle_wide*, const uint8_t*, float*);
int mrp_sample_actions_wide(mrp_handle_wide*, uint64_t, float*, void*);
int mrp_get_state_wide(mrp_handle_wide*, int32_t, int32_t, uint32_t*);
int mrp_set_state_wide(mrp_handle_wide*, int32_t, int32_t, const uint32_t*);
int mrp_enable_terminal_info_wide(mrp_handle_wide*, mrp_terminal_buffers*);
int mrp_enable_curriculum_wide(mrp_handle_wide*, double**, double**);
int mrp_set_params_wide(mrp_handle_wide*, const mrp_params*);
int mrp_get_params_wide(mrp_handle_wide*, mrp_params*);
int mrp_get_stats_wide(mrp_handle_wide*, double*, int32_t);
int64_t mrp_launch_count_wide(mrp_handle_wide*);
}
// forward a call on a fronting handle to the wide build, carrying its error text over
#define FWD(call)                                                                          \
    if (h && h->wide) {                                                                    \
        const int rc_ = (call);                                                            \
        if (rc_) snprintf(g_err, sizeof(g_err), "%s", mrp_last_error_wide());              \
        return rc_;                                                                        \
    }
#else
#define FWD(call)
#endif

extern "C" {

const char* MRP_API(mrp_last_error)(void) { return g_err; }

const char* MRP_API(mrp_backend)(void) {
#ifdef MRP_HOST_EMU
    return "host-emu (test only)";
#else
    return "cuda-sm_100a";
#endif
}

int MRP_API(mrp_set_timing)(mrp_handle* h, int32_t enable);

int MRP_API(mrp_destroy)(mrp_handle* h) {
    if (!h) return 0;
#ifndef MRP_WIDE
    if (h->wide) {
        const int rc = mrp_destroy_wide(h->wide);
        delete h;
        return rc;
    }
#endif
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    if (h->timing) MRP_API(mrp_set_timing)(h, 0);
    if (h->gx_step) cudaGraphExecDestroy(h->gx_step);
    if (h->gx_host) cudaGraphExecDestroy(h->gx_host);
    if (h->cfork) {
        for (int c = 0; c < kMaxChunks; ++c) { cudaStreamDestroy(h->cstream[c]); cudaEventDestroy(h->cjoin[c]); cudaEventDestroy(h->cpost[c]); cudaEventDestroy(h->cd2h[c]); cudaEventDestroy(h->cdone[c]); }
        cudaStreamDestroy(h->copy_stream);
        cudaStreamDestroy(h->rstream);
        if (h->refill_seen) cudaFreeHost(h->refill_seen);
        cudaEventDestroy(h->crf0);
        cudaEventDestroy(h->crf1);
        if (h->tr_init) for (int i = 0; i < 32; ++i) cudaEventDestroy(h->tr[i]);
        cudaEventDestroy(h->cfork);
        cudaEventDestroy(h->cact);
        cudaEventDestroy(h->cact0);
        cudaEventDestroy(h->cpre);
        cudaEventDestroy(h->cfree);
        cudaEventDestroy(h->cbig);
        for (int w = 0; w < kMaxWaves; ++w) { cudaEventDestroy(h->wpre[w]); cudaEventDestroy(h->wbig[w]); cudaEventDestroy(h->wjoin[w]); }
    }
#else
    free(h->emu_sm);
#endif
    DEV_FREE(h->K.S);
    DEV_FREE(h->ctab_dev);
    DEV_FREE(h->act_dev);
    DEV_FREE(h->mask_dev);
    DEV_FREE(h->K.obs);
    DEV_FREE(h->K.rew);
    DEV_FREE(h->K.done);
    DEV_FREE(h->K.trunc);
    DEV_FREE(h->K.stats);
    DEV_FREE(h->K.reset_list);
    DEV_FREE(h->K.S2);
    DEV_FREE(h->K.obs2);
    DEV_FREE(h->K.spare_ok);
    DEV_FREE(h->K.refill_list);
    DEV_FREE(h->K.refill_cnt);
    DEV_FREE(h->r_cnt); DEV_FREE(h->r_task_env); DEV_FREE(h->r_task_T); DEV_FREE(h->r_task_off); DEV_FREE(h->r_toi); DEV_FREE(h->r_post);
    DEV_FREE(h->r_narrow); DEV_FREE(h->r_pool); DEV_FREE(h->r_stats);
    DEV_FREE(h->K.cnt);
    DEV_FREE(h->K.pool);
    DEV_FREE(h->K.task_env);
    DEV_FREE(h->K.task_T);
    DEV_FREE(h->K.task_off);
    DEV_FREE(h->K.toi_list);
    DEV_FREE(h->K.post_list);
    DEV_FREE(h->K.narrow_list);
    DEV_FREE((void*)h->K.eps_env);
    DEV_FREE((void*)h->K.decay_env);
    DEV_FREE(h->K.term_obs);
    DEV_FREE(h->K.term_ret);
    DEV_FREE(h->K.term_len);
    delete h;
    return 0;
}

int MRP_API(mrp_create)(const mrp_config* cfg, mrp_handle** out) {
    if (!cfg || !out) return fail(-1, "mrp_create: null argument");
    *out = nullptr;
    if (cfg->num_envs <= 0) return fail(-2, "mrp_create: num_envs must be > 0");
    if (cfg->num_envs > (int)(0xffffffffu / (uint32_t)kMaxC)) return fail(-2, "mrp_create: num_envs too large for one handle (2^32 / contact capacity)");
#ifndef MRP_HOST_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(-3, "mrp_create: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(cudaGetLastError()));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(-4, "mrp_create: bad device ordinal");
    if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(-4, "mrp_create: cudaSetDevice failed: %s", dev_err());
#endif
    mrp_handle* h = new (std::nothrow) mrp_handle();
    if (!h) return fail(-5, "mrp_create: out of host memory");
    memset(h, 0, sizeof(*h));
#ifndef MRP_WIDE
    if ((cfg->variant >= 2 && cfg->n_agents > 2) || cfg->variant == MRP_VARIANT_SQUARE_V2) {  // MultiRobotPuzzle2(num_agents > 2), square variant: contact capacity 192 build
        const int rc = mrp_create_wide(cfg, &h->wide);
        if (rc) {
            snprintf(g_err, sizeof(g_err), "%s", mrp_last_error_wide());
            delete h;
            return rc;
        }
        *out = h;
        return 0;
    }
#endif
    float ctab[CT_WORDS];
    if (build_variant(cfg->variant, cfg->n_agents, &h->K, ctab, &h->L) != 0) {
        delete h;
        return fail(-6, "mrp_create: bad variant / n_agents (1..8 robots)");
    }
    SimConst& K = h->K;
    if (cfg->max_episode_steps > 0) { K.max_steps = cfg->max_episode_steps; h->L.max_episode_steps = cfg->max_episode_steps; }
    K.auto_reset = cfg->auto_reset ? 1 : 0;
    h->fused = getenv("MRP_FUSED_STEP") ? 1 : 0;
    h->reset_lanes = getenv("MRP_RESET_LANES") ? atoi(getenv("MRP_RESET_LANES")) : 32;
    if (h->reset_lanes < 1 || h->reset_lanes > 32 || (h->reset_lanes & (h->reset_lanes - 1))) h->reset_lanes = 32;
    h->big_ctas = getenv("MRP_BIG_CTAS") ? atoi(getenv("MRP_BIG_CTAS")) : 1;
    if (h->big_ctas < 1 || h->big_ctas > 4) h->big_ctas = 1;
    h->solver_ctas = getenv("MRP_SOLVER_CTAS") ? atoi(getenv("MRP_SOLVER_CTAS")) : 4;
    if (h->solver_ctas < 1) h->solver_ctas = 1;
    // small batches are one chunk unless the environment variables say otherwise (tests exercise chunking that way)
    auto chunks_from = [&](const char* var, int dflt) {
        int v = getenv(var) ? atoi(getenv(var)) : (cfg->num_envs >= 16384 * dflt ? dflt : 1);
        return v < 1 ? 1 : (v > kMaxChunks ? kMaxChunks : v);
    };
    h->nchunks = chunks_from("MRP_CHUNKS", 1);
    h->nchunks_host = chunks_from("MRP_CHUNKS_HOST", 8);
    K.seed = cfg->seed;
    K.env_id_base = cfg->env_id_base;
    K.N = cfg->num_envs;
    K.env0 = 0;
    K.nloc = cfg->num_envs;
    h->device = cfg->device;
    h->num_sms = 148;
#ifndef MRP_HOST_EMU
    if (cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess || h->num_sms < 1) h->num_sms = 148;
#endif
    const size_t N = (size_t)cfg->num_envs;
    int rc = 0;
    const size_t ntiles = (N + kTile - 1) / kTile;   // state tiles of 32 envs (mrp_sim.cuh)
    rc |= DEV_ALLOC(K.S, sizeof(uint32_t) * ntiles * kTile * K.w_total);
    rc |= DEV_ALLOC(h->ctab_dev, sizeof(float) * CT_WORDS);
    rc |= DEV_ALLOC(h->act_dev, sizeof(float) * N * K.act_dim);
    rc |= DEV_ALLOC(K.obs, sizeof(float) * N * K.obs_dim);
    rc |= DEV_ALLOC(K.rew, sizeof(float) * N);
    rc |= DEV_ALLOC(K.done, N);
    rc |= DEV_ALLOC(K.trunc, N);
    rc |= DEV_ALLOC(K.stats, sizeof(double) * MRP_N_STATS);
    rc |= DEV_ALLOC(K.reset_list, sizeof(int32_t) * N);
    rc |= DEV_ALLOC(K.cnt, sizeof(int32_t) * 32 * (kMaxChunks + 1 + kMaxWaves));
    // worst case: every contact slot of every env touching (never reached; pages stay untouched otherwise)
    rc |= DEV_ALLOC_RAW(K.pool, sizeof(float) * N * K.maxc * VC_WORDS);
    rc |= DEV_ALLOC(K.task_env, sizeof(int32_t) * kTaskClasses * N * K.nb);  // classes x at most one island per dynamic body
    rc |= DEV_ALLOC(K.task_T, sizeof(int32_t) * kTaskClasses * N * K.nb);
    rc |= DEV_ALLOC(K.task_off, sizeof(int32_t) * kTaskClasses * N * K.nb);
    rc |= DEV_ALLOC(K.toi_list, sizeof(int32_t) * N);
    rc |= DEV_ALLOC(K.post_list, sizeof(int32_t) * N);
    rc |= DEV_ALLOC_RAW(K.narrow_list, sizeof(uint32_t) * N * K.maxc);
    if (rc) {
        fail(-7, "mrp_create: device allocation failed: %s", dev_err());
        MRP_API(mrp_destroy)(h);
        return -7;
    }
    K.ctab = h->ctab_dev;
    K.act = h->act_dev;
    H2D(h->ctab_dev, ctab, sizeof(ctab));
    memcpy(h->ctab_host, ctab, sizeof(ctab));
    // episode counter starts at -1 so the first reset spawns episode 0; v0 goal is fixed
    {
        union { double d; uint32_t u[2]; } gx, gy;
        gx.d = K.goal_x0; gy.d = K.goal_y0;
        const int words[5] = {W_EPISODE, W_GOAL, W_GOAL + 1, W_GOAL + 2, W_GOAL + 3};
        const uint32_t vals[5] = {0xffffffffu, gx.u[0], gx.u[1], gy.u[0], gy.u[1]};
        uint32_t* row = (uint32_t*)malloc(sizeof(uint32_t) * ntiles * kTile);
        for (int w = 0; w < 5; ++w) {   // word w of every env: one 128-byte run per tile
            for (size_t i = 0; i < ntiles * kTile; ++i) row[i] = vals[w];
#ifndef MRP_HOST_EMU
            cudaMemcpy2D(K.S + (size_t)words[w] * kTile, sizeof(uint32_t) * kTile * K.w_total, row, sizeof(uint32_t) * kTile,
                         sizeof(uint32_t) * kTile, ntiles, cudaMemcpyHostToDevice);
#else
            for (size_t t = 0; t < ntiles; ++t) memcpy(K.S + t * kTile * K.w_total + (size_t)words[w] * kTile, row, sizeof(uint32_t) * kTile);
#endif
        }
        free(row);
    }
    // spare episodes (see SimConst::S2).  Not in the wide build: its fused paths borrow the env's slice of the task pool as
    // scratch, which the solver kernels of the running step are using.
#ifdef MRP_WIDE
    h->use_spares = 0;
#else
    // default from 131,072 envs: below, a step is short enough (0.6 ms at 65,536 v0 envs) for the three extra stream operations of
    // the spare list to cost 5 % (measured), and the fused respawn's tail hides less often behind nothing
    h->use_spares = getenv("MRP_SPARES") ? atoi(getenv("MRP_SPARES")) : (cfg->num_envs >= 131072 ? 1 : 0);
#endif
    if (h->use_spares && K.auto_reset) {
        const size_t sbytes = sizeof(uint32_t) * ntiles * kTile * K.w_total;
        int rs = DEV_ALLOC_RAW(K.S2, sbytes);
        rs |= DEV_ALLOC(K.obs2, sizeof(float) * N * K.obs_dim);
        rs |= DEV_ALLOC(K.spare_ok, (N + 15) / 16 * 16);
        // at most an eighth of the batch (4,096 at least) gets its spare per step: after a reset of everything the spares arrive
        // over eight steps, and an env that finishes before its spare exists takes the fused respawn
        size_t cap = ((N / 8 + kBlock - 1) / kBlock) * kBlock;
        if (cap < 4096) cap = 4096;
        if (cap > N) cap = N;
        if (getenv("MRP_REFILL_CAP") && atoi(getenv("MRP_REFILL_CAP")) > 0) cap = (size_t)atoi(getenv("MRP_REFILL_CAP")) < N ? (size_t)atoi(getenv("MRP_REFILL_CAP")) : N;
        K.refill_cap = (int32_t)cap;
        rs |= DEV_ALLOC(K.refill_list, sizeof(int32_t) * cap);
        rs |= DEV_ALLOC(K.refill_cnt, sizeof(int32_t) * 4);
        rs |= DEV_ALLOC(h->r_cnt, sizeof(int32_t) * 32);
        rs |= DEV_ALLOC_RAW(h->r_pool, sizeof(float) * cap * K.maxc * VC_WORDS);
        rs |= DEV_ALLOC(h->r_task_env, sizeof(int32_t) * kTaskClasses * cap * K.nb);
        rs |= DEV_ALLOC(h->r_task_T, sizeof(int32_t) * kTaskClasses * cap * K.nb);
        rs |= DEV_ALLOC(h->r_task_off, sizeof(int32_t) * kTaskClasses * cap * K.nb);
        rs |= DEV_ALLOC(h->r_toi, sizeof(int32_t) * cap);
        rs |= DEV_ALLOC(h->r_post, sizeof(int32_t) * cap);
        rs |= DEV_ALLOC_RAW(h->r_narrow, sizeof(uint32_t) * cap * K.maxc);
        rs |= DEV_ALLOC(h->r_stats, sizeof(double) * MRP_N_STATS);
        if (rs) {
            fail(-7, "mrp_create: device allocation failed (spare episodes): %s", dev_err());
            MRP_API(mrp_destroy)(h);
            return -7;
        }
#ifndef MRP_HOST_EMU
        cudaMemcpy(K.S2, K.S, sbytes, cudaMemcpyDeviceToDevice);   // the constant header words (v0 goal)
#else
        memcpy(K.S2, K.S, sbytes);
#endif
    }
    h->smem_bytes = sizeof(float) * ((size_t)kCtPad + (size_t)K.smem_words * kBlock);
    h->smem_broad = sizeof(float) * ((size_t)kCtPad + (size_t)(10 * K.nb + 4 * K.ndynfix) * kBlock);
    h->smem_post = sizeof(float) * ((size_t)kCtPad + (size_t)(11 * K.nb + 4 * K.ndynfix) * kBlock);
    h->smem_pre = sizeof(float) * ((size_t)kCtPad + (size_t)(8 * K.nb + 24) * kBlock);
    h->smem_vel = sizeof(float) * (size_t)(6 * (K.nb + 4)) * kBlock;
    h->smem_pos = sizeof(float) * (size_t)(9 * K.nb + 24) * kBlock;
    h->smem_big = sizeof(float) * ((size_t)(9 * K.nb + 24) + kBigRecWords) * kBigLanes;
    // the square variant's blocks lean on each other: islands with more than two contacts are the rule there, not the 1.7 % tail the
    // side kernel was built for (measured: 14.2 ms with the bulk kernels, 30.8 ms with k_solve_big, 524,288 envs)
    h->big_split = getenv("MRP_BIG") ? atoi(getenv("MRP_BIG")) : (cfg->num_envs >= 32768 && cfg->variant != MRP_VARIANT_SQUARE_V2 ? 1 : 0);
#ifndef MRP_HOST_EMU
    cudaFuncSetAttribute(k_broad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_broad);
    cudaFuncSetAttribute(k_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_pre);
    cudaFuncSetAttribute(k_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_post);
    cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    cudaFuncSetAttribute(k_reset_list, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    cudaFuncSetAttribute(k_spawn_list, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    cudaFuncSetAttribute(k_post_events, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    // experiment knobs: shared-memory carve-out (percent of the SM's 228 KB) of the per-env kernels — what is not carved out
    // is L1, which backs the lanes' local arrays (contact words, island order, actions)
    if (getenv("MRP_CARVEOUT_PRE")) cudaFuncSetAttribute(k_pre, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("MRP_CARVEOUT_PRE")));
    if (getenv("MRP_CARVEOUT_POST")) {
        cudaFuncSetAttribute(k_post, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("MRP_CARVEOUT_POST")));
    }
    if (getenv("MRP_CARVEOUT_BROAD")) cudaFuncSetAttribute(k_broad, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("MRP_CARVEOUT_BROAD")));
    cudaFuncSetAttribute(k_solve_vel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_vel);
    cudaFuncSetAttribute(k_solve_pos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_pos);
    cudaFuncSetAttribute(k_solve_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_big);
    cudaFuncSetAttribute(k_reset_mask, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    {
        // earlier chunks get higher stream priority: their CTAs are scheduled first, so they finish first and their
        // D2H copies run under the later chunks' kernels
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = least (numerically largest), hi = greatest
        const int use_prio = getenv("MRP_STREAM_PRIO") ? atoi(getenv("MRP_STREAM_PRIO")) : 1;
        for (int c = 0; c < kMaxChunks; ++c) {
            int pr = hi + c;
            if (pr > lo || !use_prio) pr = lo;
            cudaStreamCreateWithPriority(&h->cstream[c], cudaStreamNonBlocking, pr);
            cudaEventCreateWithFlags(&h->cjoin[c], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&h->cpost[c], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&h->cd2h[c], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&h->cdone[c], cudaEventDisableTiming);
        }
        cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, hi);
        // highest priority: the pass is a chain of a dozen small dependent kernels; at the priority of the step's own (large) grids
        // each of them was dispatched only when one of those had drained, so the chain ended with the step and the auto-reset
        // waited for it (+0.25 ms per step); its grids are sized to the list, so it takes little from the step
        cudaStreamCreateWithPriority(&h->rstream, cudaStreamNonBlocking, getenv("MRP_REFILL_PRIO") ? atoi(getenv("MRP_REFILL_PRIO")) : hi);
        if (cudaHostAlloc((void**)&h->refill_seen, sizeof(int32_t) * 2, cudaHostAllocDefault) == cudaSuccess) h->refill_seen[0] = (int32_t)cfg->num_envs;
        else h->refill_seen = nullptr;
        cudaEventCreateWithFlags(&h->crf0, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->crf1, cudaEventDisableTiming);
        h->host_early_copy = getenv("MRP_HOST_EARLY_COPY") ? atoi(getenv("MRP_HOST_EARLY_COPY")) : 1;
    }
    cudaEventCreateWithFlags(&h->cfork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->cact, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->cact0, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->cpre, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->cfree, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->cbig, cudaEventDisableTiming);
    for (int w = 0; w < kMaxWaves; ++w) {
        cudaEventCreateWithFlags(&h->wpre[w], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->wbig[w], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->wjoin[w], cudaEventDisableTiming);
    }
    // measured end to end (1M Heavy-v0 envs, pinned buffers): 7.09 / 6.66 / 6.25 / 6.37 ms per step with 1 / 2 / 3 / 4 waves; 524,288 envs
    // 3.91 / 3.65 / 3.58 / 3.60; v0 4.08 / 3.75 / 3.63 / 3.67 (profiles/r2_e2e_waves.md)
    // (the square variant's step is dominated by the velocity solve of its big islands, which concurrent waves only slow down:
    // 16.3 ms with one wave, 19.7 with three, 524,288 envs)
    h->host_waves = getenv("MRP_HOST_WAVES") ? atoi(getenv("MRP_HOST_WAVES")) : (cfg->num_envs >= 262144 && cfg->variant != MRP_VARIANT_SQUARE_V2 ? 3 : 1);
    if (h->host_waves < 1 || h->host_waves > kMaxWaves || h->nchunks_host < 2 * h->host_waves) h->host_waves = 1;
    {
        // wave boundaries in eighths of the chunks; every wave spans at least two chunks (its front stream and the stream of its
        // big-island kernel are the streams of its first two chunks).  MRP_HOST_WAVE_BOUNDS="3,6" overrides the inner boundaries.
        static const int dflt[kMaxWaves + 1][kMaxWaves + 1] = {{0, 0, 0, 0, 0}, {0, 8, 8, 8, 8}, {0, 3, 8, 8, 8}, {0, 3, 6, 8, 8}, {0, 2, 4, 6, 8}};
        for (int w = 0; w <= kMaxWaves; ++w) h->wave_bound[w] = dflt[h->host_waves][w] * h->nchunks_host / 8;
        if (const char* b = getenv("MRP_HOST_WAVE_BOUNDS")) {
            int v[3] = {0, 0, 0};
            const int got = sscanf(b, "%d,%d,%d", &v[0], &v[1], &v[2]);
            bool ok = got == h->host_waves - 1;
            for (int i = 0; ok && i < got; ++i) ok = v[i] >= (i ? v[i - 1] : 0) + 2 && v[i] <= h->nchunks_host - 2 * (got - i);
            for (int i = 0; ok && i < got; ++i) h->wave_bound[i + 1] = v[i];
        }
        h->wave_bound[h->host_waves] = h->nchunks_host;
    }
    {
        // measured at 1M envs: Heavy-v0 -4 %, v0 -1.5 %, v2 with 5 robots -5 %; v2 with its default two robots +9 % (hardly
        // any env owns a solver task there, so the split only adds launches): off for that case
        const bool sparse_contacts = cfg->variant >= MRP_VARIANT_V2 && h->L.n_agents <= 2;
        h->overlap_post = getenv("MRP_OVERLAP_POST") ? atoi(getenv("MRP_OVERLAP_POST")) : (cfg->num_envs >= 65536 && !sparse_contacts ? 1 : 0);
    }
    K.stage_rows = getenv("MRP_STAGE_ROWS") ? atoi(getenv("MRP_STAGE_ROWS")) : 1;
    h->use_graph = getenv("MRP_GRAPH") ? atoi(getenv("MRP_GRAPH")) : (cfg->num_envs < 32768 ? 1 : 0);
    if (check_launch("mrp_create")) { MRP_API(mrp_destroy)(h); return -10; }
#else
    h->emu_sm = (float*)calloc((size_t)K.smem_words + 8, sizeof(float));
    memcpy(h->ctab_dev, ctab, sizeof(ctab));
#endif
    *out = h;
    return 0;
}

int MRP_API(mrp_get_layout)(mrp_handle* h, mrp_layout* out) {
    FWD(mrp_get_layout_wide(h->wide, out))
    if (!h || !out) return fail(-1, "mrp_get_layout: null argument");
    *out = h->L;
    return 0;
}

int MRP_API(mrp_get_buffers)(mrp_handle* h, mrp_buffers* out) {
    FWD(mrp_get_buffers_wide(h->wide, out))
    if (!h || !out) return fail(-1, "mrp_get_buffers: null argument");
    out->action_dev = h->act_dev;
    out->obs_dev = h->K.obs;
    out->reward_dev = h->K.rew;
    out->done_dev = h->K.done;
    out->trunc_dev = h->K.trunc;
    out->stats_dev = h->K.stats;
    out->num_envs = (int32_t)h->K.N;
    out->obs_dim = h->K.obs_dim;
    out->act_dim = h->K.act_dim;
    out->reserved = 0;
    return 0;
}

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

int MRP_API(mrp_reset)(mrp_handle* h, const uint8_t* mask_dev, void* stream) {
    FWD(mrp_reset_wide(h->wide, mask_dev, stream))
    if (!h) return fail(-1, "mrp_reset: null handle");
    SimConst K = h->K;
    K.reset_mask = mask_dev;
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    k_reset_mask<<<grid_for(K.N, kBlock), kBlock, h->smem_bytes, (cudaStream_t)stream>>>(K);
    h->launches += 1;
    return check_launch("mrp_reset");
#else
    (void)stream;
    for (int64_t e = 0; e < K.N; ++e)
        if (!mask_dev || mask_dev[e]) { reset_lane(K, h->emu_sm, h->ctab_dev, e); if (K.spare_ok) K.spare_ok[e] = 0; }
    return 0;
#endif
}

#ifndef MRP_HOST_EMU
static void drain_timing(mrp_handle* h) {
    for (int i = 0; i < h->ev_n; ++i) {
        float ms = 0.0f;
        cudaEventSynchronize(h->ev1[i]);
        if (cudaEventElapsedTime(&ms, h->ev0[i], h->ev1[i]) == cudaSuccess) { h->k_ms += ms; h->k_count += 1; }
        if (!h->fused) {
            cudaEvent_t b[6] = {h->ev0[i], h->evk[i][0], h->evk[i][1], h->evk[i][2], h->evk[i][3], h->ev1[i]};
            for (int k = 0; k < 5; ++k)
                if (cudaEventElapsedTime(&ms, b[k], b[k + 1]) == cudaSuccess) h->phase_ms[k] += ms;
        }
    }
    h->ev_n = 0;
}
#endif

int MRP_API(mrp_set_timing)(mrp_handle* h, int32_t enable) {
    FWD(mrp_set_timing_wide(h->wide, enable))
    if (!h) return fail(-1, "mrp_set_timing: null handle");
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    if (enable && !h->timing)
        for (int i = 0; i < 64; ++i) {
            cudaEventCreate(&h->ev0[i]); cudaEventCreate(&h->ev1[i]);
            for (int k = 0; k < 4; ++k) cudaEventCreate(&h->evk[i][k]);
        }
    if (!enable && h->timing) {
        drain_timing(h);
        for (int i = 0; i < 64; ++i) {
            cudaEventDestroy(h->ev0[i]); cudaEventDestroy(h->ev1[i]);
            for (int k = 0; k < 4; ++k) cudaEventDestroy(h->evk[i][k]);
        }
    }
#endif
    h->timing = enable ? 1 : 0;
    return 0;
}

int MRP_API(mrp_get_timing)(mrp_handle* h, double* total_ms, int64_t* count, int32_t reset_after) {
    FWD(mrp_get_timing_wide(h->wide, total_ms, count, reset_after))
    if (!h || !total_ms || !count) return fail(-1, "mrp_get_timing: null argument");
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    if (h->timing) drain_timing(h);
#endif
    *total_ms = h->k_ms;
    *count = h->k_count;
    if (reset_after) { h->k_ms = 0.0; h->k_count = 0; }
    return 0;
}

int MRP_API(mrp_get_phase_timing)(mrp_handle* h, double* ms5, int32_t reset_after) {
    FWD(mrp_get_phase_timing_wide(h->wide, ms5, reset_after))
    if (!h || !ms5) return fail(-1, "mrp_get_phase_timing: null argument");
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    if (h->timing) drain_timing(h);
    for (int k = 0; k < 5; ++k) ms5[k] = h->phase_ms[k];
    if (reset_after) for (int k = 0; k < 5; ++k) h->phase_ms[k] = 0.0;
#else
    for (int k = 0; k < 5; ++k) ms5[k] = 0.0;
#endif
    return 0;
}

// The env range of a handle is processed as `nch` independent chunks.  Every chunk has its own counters, task
// queues, pool region and event lists (disjoint slices of the handle's arrays), so the pipelines of different chunks
// can run concurrently on separate streams: the serial tails of one chunk's persistent solver kernels and the
// host copies of mrp_step_host overlap with the other chunks' kernels.  Results do not depend on the chunking
// (envs are independent; tested in tests/test_emu_parity.py and tests/test_gpu_parity.py).
static SimConst chunk_const(const mrp_handle* h, const SimConst& K0, int c, int nch) {
    SimConst K = K0;
    const int64_t per = ((K0.N + nch - 1) / nch + kBlock - 1) / kBlock * kBlock;
    int64_t b = per * c, e = b + per;
    if (b > K0.N) b = K0.N;
    if (e > K0.N) e = K0.N;
    K.env0 = b;
    K.nloc = (int32_t)(e - b);
    K.cnt = K0.cnt + 32 * c;
    K.pool = K0.pool + (size_t)b * K0.maxc * VC_WORDS;
    K.task_env = K0.task_env + (size_t)kTaskClasses * b * K0.nb;
    K.task_T = K0.task_T + (size_t)kTaskClasses * b * K0.nb;
    K.task_off = K0.task_off + (size_t)kTaskClasses * b * K0.nb;
    K.toi_list = K0.toi_list + b;
    K.narrow_list = K0.narrow_list + (size_t)b * K0.maxc;
    K.reset_list = K0.reset_list + b;
    K.post_list = K0.post_list + b;
    (void)h;
    return K;
}

// constants of the refill pass: the handle's, re-pointed at the spare buffers and at the pass's own queues
static SimConst refill_const(const mrp_handle* h) {
    SimConst K2 = h->K;
    K2.S = h->K.S2;
    K2.obs = h->K.obs2;
    K2.stats = h->r_stats;
    K2.hidden = 1;
    K2.auto_reset = 0;
    K2.big_split = 0;
    K2.term_obs = nullptr;
    K2.idx_list = h->K.refill_list;
    K2.idx_count = h->K.refill_cnt;
    K2.env0 = 0;
    K2.nloc = h->K.refill_cap;
    K2.cnt = h->r_cnt;
    K2.pool = h->r_pool;
    K2.task_env = h->r_task_env;
    K2.task_T = h->r_task_T;
    K2.task_off = h->r_task_off;
    K2.toi_list = h->r_toi;
    K2.post_list = h->r_post;
    K2.narrow_list = h->r_narrow;
    return K2;
}
#ifndef MRP_HOST_EMU
// One chunk's phase pipeline on one stream, in two halves: the front (collide, constraint setup, solvers) and the back
// (k_post, TOI events, auto-reset), which may use a different chunking (mrp_step_host: front over the whole batch,
// back in chunks so that each chunk's D2H runs under the next chunk's kernels).  `timed` records the phase-boundary
// events (single-chunk steps only); `actions_ready` is awaited before the first kernel that reads actions (k_pre).
// Spare episodes: at the start of a step, list the envs without a valid spare, spawn their next episode in the spare buffers and
// run its hidden step as a pipeline pass over that list — all on its own stream, beside everything else of the step;
// join_refill orders a stream behind it (before auto-resets).
static void launch_refill(mrp_handle* h, cudaStream_t after) {
    if (!h->K.S2) return;
    const SimConst K2 = refill_const(h);
    // the list is made on the caller's stream (a few microseconds); only the pass itself forks off
    cudaMemsetAsync(h->K.refill_cnt, 0, sizeof(int32_t), after);
    k_refill_collect<<<grid_for((h->K.N + 15) / 16, 256), 256, 0, after>>>(h->K);
    h->launches += 1;
    // The pass's grids are sized from the list length an earlier step saw (copied to pinned memory without waiting for it), with a
    // margin: run with the capacity's grids, the pass's mostly empty CTAs — the persistent solver CTAs above all, which stay for
    // the ~1 ms tail of the slowest fresh spawn — take shared memory from the step's own kernels.  Envs beyond the grid keep their
    // flag and are listed again in the next step.
    const int64_t seen = h->refill_seen ? (int64_t)h->refill_seen[0] : (int64_t)K2.nloc;
    int64_t est = 2 * seen + 512;
    if (est > K2.nloc) est = K2.nloc;
    if (h->refill_seen) cudaMemcpyAsync(h->refill_seen, h->K.refill_cnt, sizeof(int32_t), cudaMemcpyDeviceToHost, after);
    // The pass runs only when enough spares are missing (64, or one per 16,384 envs): a dozen small kernels beside the step cost
    // 0.2-0.3 ms of a 4.9 ms step however short the list, so a few missing spares are left to accumulate — an env that finishes
    // without one takes the fused respawn, as all of them did before.
    const int64_t threshold = getenv("MRP_REFILL_MIN") ? atoi(getenv("MRP_REFILL_MIN")) : (h->K.N / 16384 > 64 ? h->K.N / 16384 : 64);
    h->refill_pending = seen >= threshold;
    if (!h->refill_pending) return;
    cudaStream_t st = h->rstream;
    cudaEventRecord(h->crf0, after);
    cudaStreamWaitEvent(st, h->crf0, 0);
    const bool trace = h->tr_init && getenv("MRP_TRACE") != nullptr;
    if (trace) cudaEventRecord(h->tr[10], st);
    const unsigned grid = grid_for(est, kBlock), nsm = (unsigned)h->num_sms;
    const unsigned sgrid = grid < nsm * (unsigned)h->solver_ctas ? grid : nsm * (unsigned)h->solver_ctas;
    const unsigned pgrid = grid < nsm * 8u ? grid : nsm * 8u;
    k_clear<<<1, 32, 0, st>>>(K2.cnt);
    k_spawn_list<<<grid, kBlock, h->smem_bytes, st>>>(h->K, K2);
    k_broad<<<grid, kBlock, h->smem_broad, st>>>(K2);
    k_narrow<<<grid < nsm * 16u ? grid : nsm * 16u, kBlock, sizeof(float) * kCtPad, st>>>(K2);
    k_pre<<<grid, kBlock, h->smem_pre, st>>>(K2, 0, K2.nloc);
    k_solve_vel<<<sgrid, kBlock, h->smem_vel, st>>>(K2);
    k_solve_pos<<<sgrid, kBlock, h->smem_pos, st>>>(K2);
    k_post<<<grid, kBlock, h->smem_post, st>>>(K2, 2);
    k_post_events<<<pgrid, kBlock, h->smem_bytes, st>>>(K2, 0);
    if (trace) cudaEventRecord(h->tr[11], st);
    cudaEventRecord(h->crf1, st);
    h->launches += 9;
}
static void join_refill(mrp_handle* h, cudaStream_t st) {
    if (h->K.S2 && h->refill_pending) cudaStreamWaitEvent(st, h->crf1, 0);
}
// auto-reset of the envs that finished in this step: a copy from their spare episode where there is one, else the fused respawn
static void launch_reset_list(mrp_handle* h, const SimConst& K, cudaStream_t st, unsigned pgrid) {
    join_refill(h, st);
    k_reset_list<<<pgrid, kBlock, h->smem_bytes, st>>>(K, h->reset_lanes);
}
static void launch_front(mrp_handle* h, const SimConst& K, cudaStream_t st, bool timed, cudaEvent_t actions_ready,
                         cudaEvent_t first_half_ready = nullptr, int64_t half = 0, cudaEvent_t* tr = nullptr, cudaStream_t big_stream = nullptr,
                         cudaEvent_t big_fork = nullptr, cudaEvent_t big_join = nullptr) {
    const unsigned grid = grid_for(K.nloc, kBlock);
    if (grid == 0) return;
    const unsigned sgrid = grid < (unsigned)h->num_sms * (unsigned)h->solver_ctas ? grid : (unsigned)h->num_sms * (unsigned)h->solver_ctas;
    k_clear<<<1, 32, 0, st>>>(K.cnt);
    if (timed) {
        if (h->ev_n == 64) drain_timing(h);
        cudaEventRecord(h->ev0[h->ev_n], st);
    }
    k_broad<<<grid, kBlock, h->smem_broad, st>>>(K);
    k_narrow<<<grid < (unsigned)h->num_sms * 16u ? grid : (unsigned)h->num_sms * 16u, kBlock, sizeof(float) * kCtPad, st>>>(K);
    if (tr) cudaEventRecord(tr[20], st);
    if (first_half_ready && half > 0 && half < K.nloc) {
        // the second half runs on another stream, beside the draining CTAs of the first (two launches in a row on one
        // stream would pay the partial last wave twice)
        cudaStream_t s2 = h->cstream[1];
        cudaEventRecord(h->cpre, st);
        cudaStreamWaitEvent(st, first_half_ready, 0);
        k_pre<<<grid_for(half, kBlock), kBlock, h->smem_pre, st>>>(K, 0, half);
        if (tr) cudaEventRecord(tr[21], st);
        cudaStreamWaitEvent(s2, h->cpre, 0);
        cudaStreamWaitEvent(s2, actions_ready, 0);
        k_pre<<<grid_for(K.nloc - half, kBlock), kBlock, h->smem_pre, s2>>>(K, half, K.nloc);
        cudaEventRecord(h->cfree, s2);
        cudaStreamWaitEvent(st, h->cfree, 0);
        h->launches += 1;
    } else {
        if (actions_ready) cudaStreamWaitEvent(st, actions_ready, 0);
        k_pre<<<grid, kBlock, h->smem_pre, st>>>(K, 0, K.nloc);
    }
    if (tr) cudaEventRecord(tr[22], st);
    if (timed) cudaEventRecord(h->evk[h->ev_n][0], st);
    // big islands beside the bulk solver kernels, as in launch_step (the third chunk stream is idle until the back half starts)
    const bool big_on = h->big_split && !timed;
    SimConst Ks = K;
    Ks.big_split = big_on ? 1 : 0;
    cudaStream_t sb = big_stream ? big_stream : h->cstream[2];
    cudaEvent_t efork = big_fork ? big_fork : h->cpre, ejoin = big_join ? big_join : h->cbig;
    if (big_on) {
        cudaEventRecord(efork, st);
        cudaStreamWaitEvent(sb, efork, 0);
        k_solve_big<<<(unsigned)h->num_sms * (unsigned)h->big_ctas, kBigLanes, h->smem_big, sb>>>(Ks);
        cudaEventRecord(ejoin, sb);
        h->launches += 1;
    }
    k_solve_vel<<<sgrid, kBlock, h->smem_vel, st>>>(Ks);
    if (timed) cudaEventRecord(h->evk[h->ev_n][1], st);
    k_solve_pos<<<sgrid, kBlock, h->smem_pos, st>>>(Ks);
    if (timed) cudaEventRecord(h->evk[h->ev_n][2], st);
    if (big_on) cudaStreamWaitEvent(st, ejoin, 0);
    h->launches += 6;
}
static void launch_back(mrp_handle* h, const SimConst& K, cudaStream_t st, bool timed, bool clear, cudaEvent_t after_post = nullptr) {
    const unsigned grid = grid_for(K.nloc, kBlock);
    if (grid == 0) return;
    const unsigned pgrid = grid < (unsigned)h->num_sms * 8u ? grid : (unsigned)h->num_sms * 8u;  // queue kernels: a few CTAs per SM
    if (clear) { k_clear<<<1, 32, 0, st>>>(K.cnt); h->launches += 1; }
    k_post<<<grid, kBlock, h->smem_post, st>>>(K, 2);
    if (after_post) cudaEventRecord(after_post, st);
    if (timed) cudaEventRecord(h->evk[h->ev_n][3], st);
    k_post_events<<<pgrid, kBlock, h->smem_bytes, st>>>(K, 0);
    h->launches += 2;
    if (timed) { cudaEventRecord(h->ev1[h->ev_n], st); h->ev_n += 1; }
    if (K.auto_reset) {
        launch_reset_list(h, K, st, pgrid);
        h->launches += 1;
    }
}
static void launch_events(mrp_handle* h, const SimConst& K, cudaStream_t st, unsigned pgrid, int free_group) {
    k_post_events<<<pgrid, kBlock, h->smem_bytes, st>>>(K, free_group);
}
// mrp_step's pipeline.  With `side_ok` (and overlap_post) the k_post of the envs that own no solver task, and the TOI-event
// pass of that group, run on a second, low-priority stream beside the solver kernels.
static void launch_step(mrp_handle* h, const SimConst& K, cudaStream_t st, bool timed, bool side_ok) {
    const unsigned grid = grid_for(K.nloc, kBlock);
    if (grid == 0) return;
    const unsigned nsm = (unsigned)h->num_sms;
    const unsigned pgrid = grid < nsm * 8u ? grid : nsm * 8u;
    const unsigned sgrid = grid < nsm * (unsigned)h->solver_ctas ? grid : nsm * (unsigned)h->solver_ctas;
    const bool side_on = side_ok && h->overlap_post && !timed;
    cudaStream_t side = h->cstream[kMaxChunks - 1];
    // MRP_TRACE=1: timeline of the step (events on every stream, printed after a device synchronise; debugging aid)
    const bool trace = getenv("MRP_TRACE") != nullptr;
    if (trace && !h->tr_init) { for (int i = 0; i < 32; ++i) cudaEventCreate(&h->tr[i]); h->tr_init = 1; }
    auto mark = [&](int i, cudaStream_t s_) { if (trace) cudaEventRecord(h->tr[i], s_); };
    k_clear<<<1, 32, 0, st>>>(K.cnt);
    mark(0, st);
    if (timed) {
        if (h->ev_n == 64) drain_timing(h);
        cudaEventRecord(h->ev0[h->ev_n], st);
    }
    k_broad<<<grid, kBlock, h->smem_broad, st>>>(K);
    k_narrow<<<grid < nsm * 16u ? grid : nsm * 16u, kBlock, sizeof(float) * kCtPad, st>>>(K);
    k_pre<<<grid, kBlock, h->smem_pre, st>>>(K, 0, K.nloc);
    mark(1, st);
    if (timed) cudaEventRecord(h->evk[h->ev_n][0], st);
    if (side_on) {
        cudaEventRecord(h->cpre, st);
        cudaStreamWaitEvent(side, h->cpre, 0);
        k_post<<<grid, kBlock, h->smem_post, side>>>(K, 0);
        mark(2, side);
#ifndef MRP_WIDE
        launch_events(h, K, side, pgrid, 1);
        mark(3, side);
#endif
        cudaEventRecord(h->cfree, side);
    }
    // big islands: velocity + position solve in one small kernel on the high-priority stream, beside the bulk solver kernels
    // (its CTAs are placed first, one per SM; the bulk kernels fill the rest of the SMs)
    const bool big_on = h->big_split && !timed;
    SimConst Ks = K;
    Ks.big_split = big_on ? 1 : 0;
    if (big_on) {
        cudaStream_t sb = h->cstream[0];
        if (!side_on) cudaEventRecord(h->cpre, st);
        cudaStreamWaitEvent(sb, h->cpre, 0);
        k_solve_big<<<nsm * (unsigned)h->big_ctas, kBigLanes, h->smem_big, sb>>>(Ks);
        mark(4, sb);
        cudaEventRecord(h->cbig, sb);
        h->launches += 1;
    }
    k_solve_vel<<<sgrid, kBlock, h->smem_vel, st>>>(Ks);
    mark(5, st);
    if (timed) cudaEventRecord(h->evk[h->ev_n][1], st);
    k_solve_pos<<<sgrid, kBlock, h->smem_pos, st>>>(Ks);
    mark(6, st);
    if (timed) cudaEventRecord(h->evk[h->ev_n][2], st);
    if (big_on) cudaStreamWaitEvent(st, h->cbig, 0);
    if (!side_on) {
        // one stream: every env in order (coalesced rows instead of the gathers of the two lists)
        k_post<<<grid, kBlock, h->smem_post, st>>>(K, 2);
        if (timed) cudaEventRecord(h->evk[h->ev_n][3], st);
        launch_events(h, K, st, pgrid, 0);
        if (timed) { cudaEventRecord(h->ev1[h->ev_n], st); h->ev_n += 1; }
        h->launches += 8;
        if (K.auto_reset) {
            launch_reset_list(h, K, st, pgrid);
            h->launches += 1;
        }
        return;
    }
    if (!side_on) k_post<<<grid, kBlock, h->smem_post, st>>>(K, 0);
    k_post<<<grid, kBlock, h->smem_post, st>>>(K, 1);
    if (timed) cudaEventRecord(h->evk[h->ev_n][3], st);
    launch_events(h, K, st, pgrid, 0);
    if (side_on) cudaStreamWaitEvent(st, h->cfree, 0);
#ifdef MRP_WIDE
    // wide build: the event pass borrows the env's slice of the task pool as constraint scratch (MRP_VC_SCRATCH), and k_pre
    // bump-allocates the solver records of ALL envs from the start of that pool — so the task-free group's event pass runs
    // here, after the solver kernels, never beside them
    launch_events(h, K, st, pgrid, 1);
#else
    if (!side_on) launch_events(h, K, st, pgrid, 1);
#endif
    if (timed) { cudaEventRecord(h->ev1[h->ev_n], st); h->ev_n += 1; }
    h->launches += 10;
    if (K.auto_reset) {
        launch_reset_list(h, K, st, pgrid);
        h->launches += 1;
    }
    mark(9, st);
    if (trace && side_on) {
        cudaDeviceSynchronize();
        static const char* nm[10] = {"start", "pre", "post_free", "events_free", "big", "vel", "pos", "post_busy", "events_busy", "end"};
        int32_t c[CNT_N];
        cudaMemcpy(c, K.cnt, sizeof(c), cudaMemcpyDeviceToHost);
        printf("[mrp trace] ms since start:");
        for (int i = 1; i <= 9; ++i) {
            float ms = 0.0f;
            if ((i == 4 && !big_on) || cudaEventElapsedTime(&ms, h->tr[0], h->tr[i]) != cudaSuccess) { cudaGetLastError(); continue; }
            printf(" %s %.3f", nm[i], ms);
        }
        if (h->refill_pending) {
            float a = 0.0f, b = 0.0f;
            if (cudaEventElapsedTime(&a, h->tr[0], h->tr[10]) == cudaSuccess && cudaEventElapsedTime(&b, h->tr[0], h->tr[11]) == cudaSuccess)
                printf(" refill pass %.3f .. %.3f", a, b);
            cudaGetLastError();
        }
        printf(" | envs free %d busy %d, TOI queue %d + %d (free), big islands %d, resets %d, spares to refill %d\n", c[CNT_FREE], c[CNT_BUSY], c[CNT_TOI],
               c[CNT_TOI_F], c[CNT_TASKS + 3] + c[CNT_TASKS_LIGHT + 3], c[CNT_RESET], h->refill_seen ? h->refill_seen[0] : -1);
    }
}static void launch_pipeline(mrp_handle* h, const SimConst& K, cudaStream_t st, bool timed) {
    if (h->fused) {  // MRP_FUSED_STEP=1: single fused kernel per step (debug / A-B comparison)
        const unsigned grid = grid_for(K.nloc, kBlock);
        if (grid == 0) return;
        const unsigned pgrid = grid < (unsigned)h->num_sms * 8u ? grid : (unsigned)h->num_sms * 8u;
        k_clear<<<1, 32, 0, st>>>(K.cnt);
        if (timed) {
            if (h->ev_n == 64) drain_timing(h);
            cudaEventRecord(h->ev0[h->ev_n], st);
        }
        k_step<<<grid, kBlock, h->smem_bytes, st>>>(K);
        if (timed) { cudaEventRecord(h->ev1[h->ev_n], st); h->ev_n += 1; }
        if (K.auto_reset) launch_reset_list(h, K, st, pgrid);
        h->launches += 3;
        return;
    }
    launch_step(h, K, st, timed, false);
}
#else
// spare episodes on the host build: same order of events as the device (refill pass at the start of a step call — spawn, then the
// hidden step as a pipeline pass over the list — copy at the auto-reset)
static void emu_narrow(mrp_handle* h, const SimConst& K);
static void emu_pre(mrp_handle* h, const SimConst& K, int64_t e);
static void emu_solvers(mrp_handle* h, const SimConst& K);
static void refill_emu(mrp_handle* h) {
    const SimConst& K = h->K;
    if (!K.S2) return;
    int n = 0;
    for (int64_t e = 0; e < K.N && n < K.refill_cap; ++e)
        if (!K.spare_ok[e]) K.refill_list[n++] = (int32_t)e;
    K.refill_cnt[0] = n;
    const SimConst K2 = refill_const(h);
    for (int i = 0; i < CNT_N; ++i) K2.cnt[i] = 0;
    for (int i = 0; i < n; ++i) spawn_spare_lane(K, K2, h->emu_sm, h->ctab_dev, K.refill_list[i]);
    for (int i = 0; i < n; ++i) {
        const int64_t e = slot_env(K2, i);
        const CMask need = broad_lane(K2, h->emu_sm, h->ctab_dev, e);
        const int c = cm_count(need);
        if (c) push_narrow(K2, e, need, atomic_add_i32(&K2.cnt[CNT_NARROW], c));
    }
    emu_narrow(h, K2);
    for (int i = 0; i < n; ++i) emu_pre(h, K2, K.refill_list[i]);
    emu_solvers(h, K2);
    for (int i = 0; i < n; ++i) post_lane(K2, h->emu_sm, h->ctab_dev, K.refill_list[i], false, nullptr);
    const int ntoi = K2.cnt[CNT_TOI];
    for (int i = 0; i < ntoi; ++i) {
        const int64_t env = K2.toi_list[i];
        MRP_VC_SCRATCH(K2, env);
        post_lane(K2, h->emu_sm, h->ctab_dev, env, true, vc_local);
    }
}
static void reset_emu(mrp_handle* h, const SimConst& K, int64_t env) {
    if (!reset_from_spare(K, env)) reset_lane(K, h->emu_sm, h->ctab_dev, env);
}
static void emu_narrow(mrp_handle* h, const SimConst& K) {
    const int nnarrow = K.cnt[CNT_NARROW];
    K.stats[MRP_STAT_PAIRS] += nnarrow;
    for (int i = 0; i < nnarrow; ++i) narrow_item(K, h->ctab_dev, K.narrow_list[i]);
}
static void emu_pre(mrp_handle* h, const SimConst& K, int64_t e) {
    uint32_t m12[2] = {0u, 0u};
    const int T = pre_lane(K, h->emu_sm, h->ctab_dev, e, m12);
    K.stats[MRP_STAT_M1] += m12[0]; K.stats[MRP_STAT_M2] += m12[1];
    if (T == 0) K.post_list[K.cnt[CNT_FREE]++] = (int32_t)e;
    else K.post_list[K.nloc - 1 - K.cnt[CNT_BUSY]++] = (int32_t)e;
}
static void emu_solvers(mrp_handle* h, const SimConst& K);
// mrp_step_host's front half (collide, setup, solvers over contiguous envs: k_broad, k_narrow, k_pre, solver kernels)
static void run_front_emu(mrp_handle* h, const SimConst& K) {
    for (int i = 0; i < CNT_N; ++i) K.cnt[i] = 0;
    const int64_t e0 = K.env0, e1 = K.env0 + K.nloc;
    // the same phases the device runs as kernels, executed as loops
    for (int64_t e = e0; e < e1; ++e) {
        const CMask need = broad_lane(K, h->emu_sm, h->ctab_dev, e);
        const int n = cm_count(need);
        if (n) push_narrow(K, e, need, atomic_add_i32(&K.cnt[CNT_NARROW], n));
    }
    emu_narrow(h, K);
    for (int64_t e = e0; e < e1; ++e) emu_pre(h, K, e);
    emu_solvers(h, K);
}
static void emu_solvers(mrp_handle* h, const SimConst& K0) {
    SimConst K = K0;
    K.big_split = h->big_split;
    if (K.big_split) {   // k_solve_big: velocity + position solve of the islands with more than two contacts
        static float rec[kBigRecWords];
        uint32_t flops = 0u;
        Sim s(K, h->emu_sm, h->ctab_dev, 0, nullptr, 9);
        const int ntasks = task_count(K, 3);
        for (int i = 0; i < ntasks; ++i) big_task_lane(K, s, task_slot(K, 3, i), rec, flops);
    }
    for (int cls = kTaskClasses - 1 - (K.big_split ? 1 : 0); cls >= 0; --cls) {
        const int ntasks = task_count(K, cls);
        for (int i = 0; i < ntasks; ++i) {
            Sim s(K, h->emu_sm, h->ctab_dev, 0, nullptr, 6);
            VelTask vt;
            Sim::VelReg st1;
            vel_task_begin(K, s, vt, task_slot(K, cls, i));
            float imp3[4 * kMaxC];
            float* const imp = cls == 3 ? imp3 : nullptr;
            if (cls == 2) s.vr_begin_pair(vt.st, st1); else s.vr_begin(vt.st, vt.T, imp);
            if (cls == 0) { while (vt.ops += 2, !s.vr_sweep_single<1>(vt.st, 180)) {} }
            else if (cls == 1) { while (vt.ops += 3, !s.vr_sweep_single<2>(vt.st, 180)) {} }
            else if (cls == 2) { while (vt.ops += 5, !s.vr_sweep_pair(vt.st, st1, 180)) {} }
            else { while (vt.ops += (uint32_t)vt.st.vpc + 1u, !s.vr_trip_contact(vt.st, 180, imp)) {} }
            vel_task_end(K, s, vt);
        }
    }
    const int ntasks = task_count_all(K);
    for (int i = 0; i < ntasks; ++i) {
        Sim s(K, h->emu_sm, h->ctab_dev, 0, nullptr, 9);
        PosTask pt;
        pos_task_begin(K, s, pt, task_slot_any(K, i));
        while (!s.pos_trip(pt.st, pt.T, 60, -1, -1)) {}
        pos_task_end(K, s, pt);
    }
}
static void run_back_emu(mrp_handle* h, const SimConst& K, bool clear, bool by_lists = false) {
    if (clear) for (int i = 0; i < CNT_N; ++i) K.cnt[i] = 0;
    const int64_t e0 = K.env0, e1 = K.env0 + K.nloc;
    if (by_lists) {  // the order of the device's overlapped step: envs without solver tasks first, then the others
        for (int i = 0; i < K.cnt[CNT_FREE]; ++i) post_lane(K, h->emu_sm, h->ctab_dev, K.post_list[i], false, nullptr, true);
        for (int i = 0; i < K.cnt[CNT_BUSY]; ++i) post_lane(K, h->emu_sm, h->ctab_dev, K.post_list[K.nloc - 1 - i], false, nullptr);
    } else {
        for (int64_t e = e0; e < e1; ++e) post_lane(K, h->emu_sm, h->ctab_dev, e, false, nullptr);
    }
    for (int grp = 1; grp >= 0; --grp) {   // the task-free group's queue first, as on the device
        const int ntoi = K.cnt[grp ? CNT_TOI_F : CNT_TOI];
        for (int i = 0; i < ntoi; ++i) {
            const int64_t env = K.toi_list[grp ? K.nloc - 1 - i : i];
            MRP_VC_SCRATCH(K, env);
            post_lane(K, h->emu_sm, h->ctab_dev, env, true, vc_local);
        }
    }
    if (K.auto_reset) {
        const int nreset = K.cnt[CNT_RESET];
        for (int i = 0; i < nreset; ++i) reset_emu(h, K, K.reset_list[i]);
    }
}
static void run_pipeline_emu(mrp_handle* h, const SimConst& K) {
    if (h->fused) {
        for (int i = 0; i < CNT_N; ++i) K.cnt[i] = 0;
        for (int64_t e = K.env0; e < K.env0 + K.nloc; ++e) step_lane(K, h->emu_sm, h->ctab_dev, e);
        if (K.auto_reset) {
            const int nreset = K.cnt[CNT_RESET];
            for (int i = 0; i < nreset; ++i) reset_emu(h, K, K.reset_list[i]);
        }
        return;
    }
    run_front_emu(h, K);
    run_back_emu(h, K, false, true);
}
#endif

// chunk c of the back half when the front ran over the whole batch: own counters (blocks 1..), own event / reset queues
static SimConst back_chunk_const(const mrp_handle* h, const SimConst& K0, int c, int nch) {
    SimConst K = chunk_const(h, K0, c, nch);
    K.cnt = K0.cnt + 32 * (c + 1);
    return K;
}

// front-half wave w of mrp_step_host: the envs of back chunks [c0, c1), with its own counters (blocks behind those of the back
// chunks), queues and pool slice
static SimConst wave_const(const mrp_handle* h, const SimConst& K0, int w, int c0, int c1, int nch) {
    const SimConst first = chunk_const(h, K0, c0, nch), last = chunk_const(h, K0, c1 - 1, nch);
    SimConst K = first;
    K.nloc = (int32_t)(last.env0 + last.nloc - first.env0);
    K.cnt = K0.cnt + 32 * (kMaxChunks + 1 + w);
    return K;
}

// chunks of one step call: 1 while the per-phase timers are on (their events live on one stream)
static int step_chunks(const mrp_handle* h, int wanted) {
    int nch = h->timing ? 1 : wanted;
    if ((int64_t)nch > (h->K.N + kBlock - 1) / kBlock) nch = (int)((h->K.N + kBlock - 1) / kBlock);
    return nch < 1 ? 1 : nch;
}

}  // extern "C"
#ifndef MRP_HOST_EMU
// a step qualifies for graph replay when all of its launches go to one stream and nothing is recorded between them
static bool graph_ok(const mrp_handle* h, int nch) {
    return h->use_graph && nch == 1 && !h->timing && !h->fused && !h->overlap_post && !h->big_split && !h->K.S2 && !getenv("MRP_TRACE");
}
// capture `body` (launches on the capture stream) into an executable graph; returns nullptr when capture is not possible
template <typename F>
static cudaGraphExec_t capture_graph(mrp_handle* h, cudaStream_t cs, int64_t* launches, F body) {
    const int64_t l0 = h->launches;
    if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    body(cs);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(cs, &g);
    *launches = h->launches - l0;
    h->launches = l0;   // counted per replay
    cudaGraphExec_t x = nullptr;
    if (e != cudaSuccess || !g || cudaGraphInstantiate(&x, g, 0) != cudaSuccess) x = nullptr;
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    return x;
}
static bool host_pinned(const void* p) {
    if (!p) return true;
    cudaPointerAttributes pa;
    const bool ok = cudaPointerGetAttributes(&pa, p) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return ok;
}
#endif
extern "C" {

int MRP_API(mrp_step)(mrp_handle* h, const float* actions_dev, void* stream) {
    FWD(mrp_step_wide(h->wide, actions_dev, stream))
    if (!h) return fail(-1, "mrp_step: null handle");
    h->steps_done += 1;
    SimConst K = h->K;
    if (actions_dev) K.act = actions_dev;
    const int nch = step_chunks(h, h->nchunks);
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (graph_ok(h, nch)) {
        // compared as stored in the handle (memcpy of the same object: padding bytes included and stable)
        if (h->gx_step && (memcmp(&h->gk_step, &h->K, sizeof(SimConst)) != 0 || h->gp_act != K.act)) { cudaGraphExecDestroy(h->gx_step); h->gx_step = nullptr; }
        if (!h->gx_step) {
            h->gx_step = capture_graph(h, h->cstream[0], &h->gl_step, [&](cudaStream_t cs) { launch_step(h, chunk_const(h, K, 0, 1), cs, false, false); });
            memcpy(&h->gk_step, &h->K, sizeof(SimConst));
            h->gp_act = K.act;
            if (!h->gx_step) h->use_graph = 0;   // capture unavailable: plain launches from now on
        }
    }
    if (graph_ok(h, nch) && h->gx_step) {
        h->launches += h->gl_step;
        if (cudaGraphLaunch(h->gx_step, st) != cudaSuccess) return fail(-10, "mrp_step: graph launch failed: %s", dev_err());
        return check_launch("mrp_step");
    }
    launch_refill(h, st);   // spare episodes of the envs that were reset in the previous step, beside this step's kernels
    if (nch == 1 && !h->fused) {
        launch_step(h, chunk_const(h, K, 0, 1), st, h->timing != 0, true);
    } else if (nch == 1) {
        launch_pipeline(h, chunk_const(h, K, 0, 1), st, h->timing != 0);
    } else {
        // fork the chunk pipelines off the caller's stream and join them back: stream-ordered like one kernel
        cudaEventRecord(h->cfork, st);
        for (int c = 0; c < nch; ++c) {
            cudaStreamWaitEvent(h->cstream[c], h->cfork, 0);
            launch_pipeline(h, chunk_const(h, K, c, nch), h->cstream[c], false);
            cudaEventRecord(h->cjoin[c], h->cstream[c]);
            cudaStreamWaitEvent(st, h->cjoin[c], 0);
        }
    }
    return check_launch("mrp_step");
#else
    (void)stream;
    refill_emu(h);
    for (int c = 0; c < nch; ++c) run_pipeline_emu(h, chunk_const(h, K, c, nch));
    return 0;
#endif
}

int MRP_API(mrp_step_host)(mrp_handle* h, const float* actions_host, float* obs_host, float* reward_host, uint8_t* done_host,
                  uint8_t* trunc_host) {
    FWD(mrp_step_host_wide(h->wide, actions_host, obs_host, reward_host, done_host, trunc_host))
    if (!h || !actions_host) return fail(-1, "mrp_step_host: null argument");
    h->steps_done += 1;
    const SimConst& K0 = h->K;
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    // The front half (collide, setup, solvers) runs over the whole batch on the first chunk stream — the action H2D is
    // queued beside it on the second and only k_pre waits for it.  The back half (k_post, TOI events, auto-reset) runs
    // per chunk on prioritised streams, each followed by the D2H of its result rows, so the PCIe copies (the host
    // buffers should be pinned) run under the remaining chunks' kernels.
    const int nch = step_chunks(h, h->nchunks_host);
    const size_t N = (size_t)K0.N;
    if (graph_ok(h, nch)) {
        const void* hp[5] = {actions_host, obs_host, reward_host, done_host, trunc_host};
        if (h->gx_host && (memcmp(&h->gk_host, &K0, sizeof(SimConst)) != 0 || memcmp(h->gp_host, hp, sizeof(hp)) != 0)) {
            cudaGraphExecDestroy(h->gx_host);
            h->gx_host = nullptr;
        }
        if (!h->gx_host && (memcmp(h->gp_host, hp, sizeof(hp)) != 0 || memcmp(&h->gk_host, &K0, sizeof(SimConst)) != 0)) {
            // pageable buffers are not captured (their copies are staged by the driver at call time): plain path
            bool pinned = true;
            for (const void* p : hp) pinned = pinned && host_pinned(p);
            memcpy(h->gp_host, hp, sizeof(hp));
            memcpy(&h->gk_host, &K0, sizeof(SimConst));
            if (pinned)
                h->gx_host = capture_graph(h, h->cstream[0], &h->gl_host, [&](cudaStream_t cs) {
                    cudaMemcpyAsync(h->act_dev, actions_host, sizeof(float) * N * K0.act_dim, cudaMemcpyHostToDevice, cs);
                    launch_step(h, chunk_const(h, K0, 0, 1), cs, false, false);
                    if (obs_host) cudaMemcpyAsync(obs_host, K0.obs, sizeof(float) * N * K0.obs_dim, cudaMemcpyDeviceToHost, cs);
                    if (reward_host) cudaMemcpyAsync(reward_host, K0.rew, sizeof(float) * N, cudaMemcpyDeviceToHost, cs);
                    if (done_host) cudaMemcpyAsync(done_host, K0.done, N, cudaMemcpyDeviceToHost, cs);
                    if (trunc_host) cudaMemcpyAsync(trunc_host, K0.trunc, N, cudaMemcpyDeviceToHost, cs);
                });
        }
        if (h->gx_host) {
            cudaStream_t cs = h->cstream[0];
            cudaEventRecord(h->cfork, 0);  // order after whatever the caller queued on the default stream
            cudaStreamWaitEvent(cs, h->cfork, 0);
            h->launches += h->gl_host;
            if (cudaGraphLaunch(h->gx_host, cs) != cudaSuccess) return fail(-10, "mrp_step_host: graph launch failed: %s", dev_err());
            if (cudaStreamSynchronize(cs) != cudaSuccess) return fail(-9, "mrp_step_host: %s", dev_err());
            return 0;
        }
    }
    cudaEvent_t* const tr = h->tr;
    const bool trace = getenv("MRP_TRACE") != nullptr;
    if (trace && !h->tr_init) { for (int i = 0; i < 32; ++i) cudaEventCreate(&tr[i]); h->tr_init = 1; }
    if (trace) cudaEventRecord(tr[0], 0);
    cudaEventRecord(h->cfork, 0);  // order after whatever the caller queued on the default stream
    cudaStream_t s_front = h->cstream[0], s_h2d = h->cstream[kMaxChunks - 1];
    cudaStreamWaitEvent(s_front, h->cfork, 0);
    cudaStreamWaitEvent(s_h2d, h->cfork, 0);
    launch_refill(h, s_front);   // spare episodes, beside this step's kernels
    // Front-half waves (from 262,144 envs): the env range is cut at back-chunk boundaries into waves that run collide / setup /
    // solvers on streams of falling priority, beside each other (the per-env kernels leave most issue slots idle).  The first
    // wave's rows are post-processed and on their way to the host while the later waves are still in their solver kernels, so
    // only the copies of the last wave stay exposed.
    const int waves = (h->host_waves > 1 && nch == h->nchunks_host && !h->fused) ? h->host_waves : 1;
    const int* const wb = h->wave_bound;
    auto wave_of = [&](int c) { int w = 0; while (w + 1 < waves && c >= wb[w + 1]) ++w; return w; };
    // the action rows go up in two halves: k_pre of the first half starts while the second is still in flight
    const size_t half = nch > 1 ? (waves > 1 ? (size_t)chunk_const(h, K0, wb[1], nch).env0 : (N / 2 + kBlock - 1) / kBlock * kBlock) : 0;
    const size_t row = sizeof(float) * K0.act_dim;
    if (half && cudaMemcpyAsync(h->act_dev, actions_host, row * half, cudaMemcpyHostToDevice, s_h2d) != cudaSuccess)
        return fail(-8, "mrp_step_host: H2D failed: %s", dev_err());
    if (half) cudaEventRecord(h->cact0, s_h2d);
    if (cudaMemcpyAsync(h->act_dev + half * K0.act_dim, actions_host + half * K0.act_dim, row * (N - half), cudaMemcpyHostToDevice, s_h2d) != cudaSuccess)
        return fail(-8, "mrp_step_host: H2D failed: %s", dev_err());
    cudaEventRecord(h->cact, s_h2d);
    if (trace) cudaEventRecord(tr[1], s_h2d);
    auto copy_out = [&](cudaStream_t st, size_t b, size_t n) {
        if (obs_host) cudaMemcpyAsync(obs_host + b * K0.obs_dim, K0.obs + b * K0.obs_dim, sizeof(float) * n * K0.obs_dim, cudaMemcpyDeviceToHost, st);
        if (reward_host) cudaMemcpyAsync(reward_host + b, K0.rew + b, sizeof(float) * n, cudaMemcpyDeviceToHost, st);
        if (done_host) cudaMemcpyAsync(done_host + b, K0.done + b, n, cudaMemcpyDeviceToHost, st);
        if (trunc_host) cudaMemcpyAsync(trunc_host + b, K0.trunc + b, n, cudaMemcpyDeviceToHost, st);
    };
    unsigned zc_chunks = 0;   // back chunks on the early-copy path: their reward / done / truncation values leave at the end
    if (nch == 1 || h->fused) {
        cudaStreamWaitEvent(s_front, h->cact, 0);
        launch_pipeline(h, chunk_const(h, K0, 0, 1), s_front, h->timing != 0);
        copy_out(s_front, 0, N);
    } else {
        if (waves > 1) {
            for (int w = 0; w < waves; ++w) {
                cudaStream_t sw = h->cstream[wb[w]];
                if (w) cudaStreamWaitEvent(sw, h->cfork, 0);
                launch_front(h, wave_const(h, K0, w, wb[w], wb[w + 1], nch), sw, false, w ? h->cact : h->cact0, nullptr, 0, (trace && !w) ? tr : nullptr,
                             h->cstream[wb[w] + 1], h->wpre[w], h->wbig[w]);
                cudaEventRecord(h->wjoin[w], sw);
                if (trace && !w) cudaEventRecord(tr[2], sw);
                if (trace && w == waves - 1) { cudaEventRecord(tr[21], sw); }
            }
        } else {
            launch_front(h, chunk_const(h, K0, 0, 1), s_front, false, h->cact, h->cact0, (int64_t)half, trace ? tr : nullptr);
            cudaEventRecord(h->cjoin[0], s_front);
            if (trace) cudaEventRecord(tr[2], s_front);
        }
        // pinned obs buffer the device can address: early bulk copy + zero-copy fix-up of the rewritten rows
        float* obs_zc = nullptr;
        if (obs_host && h->host_early_copy) {
            if (h->zc_host != (const void*)obs_host) {
                cudaPointerAttributes pa;
                h->zc_host = obs_host;
                h->zc_dev = nullptr;
                if (cudaPointerGetAttributes(&pa, obs_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer &&
                    ((uintptr_t)pa.devicePointer & 15) == 0)
                    h->zc_dev = (float*)pa.devicePointer;
                cudaGetLastError();
            }
            obs_zc = h->zc_dev;
        }
        for (int c = 0; c < nch; ++c) {
            const SimConst K = back_chunk_const(h, K0, c, nch);
            if (K.nloc == 0) continue;
            cudaStream_t st = h->cstream[c];
            // front of this chunk's envs: the whole batch, or its wave
            cudaEvent_t front_done = waves > 1 ? h->wjoin[wave_of(c)] : h->cjoin[0];
            if (obs_zc) {
                cudaStreamWaitEvent(st, front_done, 0);
                if (c > 0) cudaStreamWaitEvent(st, h->cpost[c - 1], 0);
                launch_back(h, K, st, false, true, h->cpost[c]);
                if (trace) cudaEventRecord(tr[3 + 2 * c], st);
                const size_t b = (size_t)K.env0, n = (size_t)K.nloc;
                cudaStreamWaitEvent(h->copy_stream, h->cpost[c], 0);
                cudaMemcpyAsync(obs_host + b * K0.obs_dim, K0.obs + b * K0.obs_dim, sizeof(float) * n * K0.obs_dim, cudaMemcpyDeviceToHost, h->copy_stream);
                cudaEventRecord(h->cd2h[c], h->copy_stream);
                cudaStreamWaitEvent(st, h->cd2h[c], 0);
                for (int which = 0; which < (K.auto_reset ? 2 : 1); ++which) {
                    if (K0.obs_dim % 4 == 0) k_out_rows<float4><<<16, 256, 0, st>>>(K, which, reinterpret_cast<float4*>(obs_zc));
                    else k_out_rows<float><<<16, 256, 0, st>>>(K, which, obs_zc);
                    h->launches += 1;
                }
                if (trace) cudaEventRecord(tr[4 + 2 * c], st);
                cudaEventRecord(h->cdone[c], st);
                zc_chunks |= 1u << c;
                continue;
            }
            // k_post of chunk c starts when k_post of chunk c-1 has finished: at that moment the (higher-priority, large
            // shared memory) TOI-event and reset kernels of chunk c-1 get the draining SMs first, and chunk c's k_post
            // fills the rest.  Launched all at once, the k_post CTAs of later chunks would keep re-occupying the SMs and
            // starve those kernels until every k_post had drained (measured: all chunks finished together).
            cudaStreamWaitEvent(st, front_done, 0);
            if (c > 0) cudaStreamWaitEvent(st, h->cpost[c - 1], 0);
            launch_back(h, K, st, false, true, h->cpost[c]);
            if (trace) cudaEventRecord(tr[3 + 2 * c], st);
            copy_out(st, (size_t)K.env0, (size_t)K.nloc);
            if (trace) cudaEventRecord(tr[4 + 2 * c], st);
        }
    }
    if (zc_chunks) {
        // reward / done / truncation of the whole batch leave last, as three copies behind every chunk's passes.  Issued per chunk,
        // between the bulk copies, they held those up: the copy engine takes its work in host issue order, so a small copy waiting
        // for chunk c's event / reset tail (~0.4 ms) kept the rows of chunk c + 1 from starting (measured: every bulk copy ended
        // within 0.2 ms of the last one, 0.5 ms late)
        cudaStream_t st = h->copy_stream;
        for (int c = 0; c < nch; ++c) if ((zc_chunks >> c) & 1u) cudaStreamWaitEvent(st, h->cdone[c], 0);
        if (reward_host) cudaMemcpyAsync(reward_host, K0.rew, sizeof(float) * N, cudaMemcpyDeviceToHost, st);
        if (done_host) cudaMemcpyAsync(done_host, K0.done, N, cudaMemcpyDeviceToHost, st);
        if (trunc_host) cudaMemcpyAsync(trunc_host, K0.trunc, N, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) return fail(-9, "mrp_step_host: %s", dev_err());
    }
    int rc = check_launch("mrp_step_host");
    for (int c = 0; c < nch; ++c)
        if (cudaStreamSynchronize(h->cstream[c]) != cudaSuccess && !rc) rc = fail(-9, "mrp_step_host: %s", dev_err());
    if (cudaStreamSynchronize(s_h2d) != cudaSuccess && !rc) rc = fail(-9, "mrp_step_host: %s", dev_err());
    if (trace && nch > 1) {
        float ms;
        cudaEventElapsedTime(&ms, tr[0], tr[1]); printf("h2d_done %.2f", ms);
        cudaEventElapsedTime(&ms, tr[0], tr[20]); printf(" narrow_done %.2f", ms);
        cudaEventElapsedTime(&ms, tr[0], tr[21]); printf(" pre0_done %.2f", ms);
        cudaEventElapsedTime(&ms, tr[0], tr[22]); printf(" pre_done %.2f", ms);
        cudaEventElapsedTime(&ms, tr[0], tr[2]); printf(" front_done %.2f", ms);
        for (int c = 0; c < nch; ++c) { cudaEventElapsedTime(&ms, tr[0], tr[3 + 2 * c]); printf(" | back%d %.2f", c, ms); cudaEventElapsedTime(&ms, tr[0], tr[4 + 2 * c]); printf(" d2h%d %.2f", c, ms); }
        printf("\n");
    }
    return rc;
#else
    const size_t N = (size_t)K0.N;
    memcpy(h->act_dev, actions_host, sizeof(float) * N * K0.act_dim);
    refill_emu(h);
    const int nch = step_chunks(h, h->nchunks_host);
    if (nch == 1 || h->fused) {
        run_pipeline_emu(h, chunk_const(h, K0, 0, 1));
    } else {  // front over the whole batch, back per chunk (as the device path)
        run_front_emu(h, chunk_const(h, K0, 0, 1));
        for (int c = 0; c < nch; ++c) run_back_emu(h, back_chunk_const(h, K0, c, nch), true);
    }
    if (obs_host) memcpy(obs_host, K0.obs, sizeof(float) * N * K0.obs_dim);
    if (reward_host) memcpy(reward_host, K0.rew, sizeof(float) * N);
    if (done_host) memcpy(done_host, K0.done, N);
    if (trunc_host) memcpy(trunc_host, K0.trunc, N);
    return 0;
#endif
}

int MRP_API(mrp_reset_host)(mrp_handle* h, const uint8_t* mask_host, float* obs_host) {
    FWD(mrp_reset_host_wide(h->wide, mask_host, obs_host))
    if (!h) return fail(-1, "mrp_reset_host: null handle");
    const size_t N = (size_t)h->K.N;
    uint8_t* mask_dev = nullptr;
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(-9, "mrp_reset_host: %s", dev_err());   // order after steps queued on any stream
#endif
    if (mask_host) {
        if (!h->mask_dev && DEV_ALLOC_RAW(h->mask_dev, N)) return fail(-7, "mrp_reset_host: device allocation failed: %s", dev_err());
        mask_dev = h->mask_dev;
        if (H2D(mask_dev, mask_host, N)) return fail(-8, "mrp_reset_host: H2D failed: %s", dev_err());
    }
    int rc = MRP_API(mrp_reset)(h, mask_dev, nullptr);
    if (rc) return rc;
    if (obs_host && D2H(obs_host, h->K.obs, sizeof(float) * N * h->K.obs_dim)) return fail(-9, "mrp_reset_host: D2H failed: %s", dev_err());
#ifndef MRP_HOST_EMU
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(-9, "mrp_reset_host: %s", dev_err());
#endif
    return 0;
}

int MRP_API(mrp_obs_v3)(mrp_handle* h, float* out_dev, void* stream) {
    if (!h || !out_dev) return fail(-1, "mrp_obs_v3: null argument");
#ifndef MRP_WIDE
    if (h->wide) return fail(-6, "mrp_obs_v3: holonomic (v0 / Heavy-v0) envs only");
#endif
    if (h->K.v2) return fail(-6, "mrp_obs_v3: holonomic (v0 / Heavy-v0) envs only");
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    k_obs_v3<<<grid_for(h->K.N, 128), 128, 0, (cudaStream_t)stream>>>(h->K, out_dev);
    h->launches += 1;
    return check_launch("mrp_obs_v3");
#else
    (void)stream;
    for (int64_t e = 0; e < h->K.N; ++e) obs_v3_lane(h->K, out_dev, e);
    return 0;
#endif
}

int MRP_API(mrp_sample_actions)(mrp_handle* h, uint64_t step_index, float* dst_dev, void* stream) {
    FWD(mrp_sample_actions_wide(h->wide, step_index, dst_dev, stream))
    if (!h) return fail(-1, "mrp_sample_actions: null handle");
    float* dst = dst_dev ? dst_dev : h->act_dev;
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    k_sample_actions<<<grid_for(h->K.N, 256), 256, 0, (cudaStream_t)stream>>>(h->K, dst, step_index);
    h->launches += 1;
    return check_launch("mrp_sample_actions");
#else
    (void)stream;
    for (int64_t e = 0; e < h->K.N; ++e) sample_actions_lane(h->K, dst, step_index, e);
    return 0;
#endif
}

// ---- canonical state records <-> internal [word][env] layout (host side) --------------
// buf is [word][count] (word-major over the requested envs); the device state is tiled [tile][word][32] (mrp_sim.cuh): whole
// tiles covering the range travel as one contiguous copy and are re-ordered on the host
static int fetch_internal(mrp_handle* h, int64_t begin, int64_t count, uint32_t* buf) {
    const SimConst& K = h->K;
    if (count == 0) return 0;
    const int64_t t0 = begin / kTile, t1 = (begin + count - 1) / kTile + 1;
    const size_t tw = (size_t)kTile * K.w_total;
    uint32_t* tiles = (uint32_t*)malloc(sizeof(uint32_t) * tw * (size_t)(t1 - t0));
    if (!tiles) return -1;
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    if (cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpy(tiles, K.S + tw * t0, sizeof(uint32_t) * tw * (size_t)(t1 - t0), cudaMemcpyDeviceToHost) != cudaSuccess) { free(tiles); return -1; }
#else
    memcpy(tiles, K.S + tw * t0, sizeof(uint32_t) * tw * (size_t)(t1 - t0));
#endif
    for (int64_t e = 0; e < count; ++e) {
        const int64_t env = begin + e;
        const uint32_t* src = tiles + (size_t)(env / kTile - t0) * tw + (env % kTile);
        for (int w = 0; w < K.w_total; ++w) buf[(size_t)w * count + e] = src[(size_t)w * kTile];
    }
    free(tiles);
    return 0;
}
static int push_internal(mrp_handle* h, int64_t begin, int64_t count, const uint32_t* buf) {
    const SimConst& K = h->K;
    if (count == 0) return 0;
    const int64_t t0 = begin / kTile, t1 = (begin + count - 1) / kTile + 1;
    const size_t tw = (size_t)kTile * K.w_total, bytes = sizeof(uint32_t) * tw * (size_t)(t1 - t0);
    uint32_t* tiles = (uint32_t*)malloc(bytes);
    if (!tiles) return -1;
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    // steps queued on any stream finish before the state is replaced; envs of the edge tiles outside the range keep their words
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(tiles, K.S + tw * t0, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { free(tiles); return -1; }
#else
    memcpy(tiles, K.S + tw * t0, bytes);
#endif
    for (int64_t e = 0; e < count; ++e) {
        const int64_t env = begin + e;
        uint32_t* dst = tiles + (size_t)(env / kTile - t0) * tw + (env % kTile);
        for (int w = 0; w < K.w_total; ++w) dst[(size_t)w * kTile] = buf[(size_t)w * count + e];
    }
#ifndef MRP_HOST_EMU
    if (cudaMemcpy(K.S + tw * t0, tiles, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { free(tiles); return -1; }
    free(tiles);
    k_fix_rot<<<grid_for(count, 128), 128>>>(K, begin, count);
    h->launches += 1;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
#else
    memcpy(K.S + tw * t0, tiles, bytes);
    free(tiles);
    for (int64_t e = 0; e < count; ++e) fix_rot_lane(K, begin + e);
#endif
    if (K.spare_ok) DEV_ZERO(K.spare_ok + begin, (size_t)count);   // the episode counters may have changed: spares are recomputed
    return 0;
}

int MRP_API(mrp_get_state)(mrp_handle* h, int32_t env_begin, int32_t env_count, uint32_t* words) {
    FWD(mrp_get_state_wide(h->wide, env_begin, env_count, words))
    if (!h || !words) return fail(-1, "mrp_get_state: null argument");
    if (env_begin < 0 || env_count < 0 || (int64_t)env_begin + env_count > h->K.N) return fail(-2, "mrp_get_state: bad env range");
    const SimConst& K = h->K;
    const mrp_layout& L = h->L;
    const size_t C = (size_t)env_count;
    uint32_t* buf = (uint32_t*)malloc(sizeof(uint32_t) * C * K.w_total);
    if (!buf) return fail(-5, "mrp_get_state: out of host memory");
    if (fetch_internal(h, env_begin, env_count, buf)) { free(buf); return fail(-9, "mrp_get_state: copy failed: %s", dev_err()); }
    auto I = [&](int w, size_t e) -> uint32_t { return buf[(size_t)w * C + e]; };
    memset(words, 0, sizeof(uint32_t) * C * L.state_words);
    for (size_t e = 0; e < C; ++e) {
        uint32_t* o = words + e * L.state_words;
        o[0] = I(W_ELAPSED, e); o[1] = I(W_EPISODE, e); o[2] = I(W_INPLACE, e);
        int nc = (int)I(W_NC, e);
        o[3] = (uint32_t)nc;
        uint32_t gc = I(W_GOALC, e);
        for (int i = 0; i < K.n; ++i) o[L.off_goal_contact + i] = (gc >> i) & 1u;
        for (int b = 0; b < K.nb; ++b)
            for (int f = 0; f < 6; ++f) o[L.off_bodies + 6 * b + f] = I(K.w_body + kBodyWords * b + f, e);
        for (int i = 0; i < 2 * (K.n + 1); ++i) o[L.off_dists + i] = I(W_DIST + i, e);
        for (int i = 0; i < 4; ++i) o[L.off_goal + i] = I(W_GOAL + i, e);
        o[L.off_episode_acc] = I(W_EPRET, e); o[L.off_episode_acc + 1] = I(W_EPRET + 1, e);
        o[L.off_episode_acc + 2] = I(W_EPLEN, e);
        for (int i = 0; i < 4 * K.ndynfix; ++i) o[L.off_aabb + i] = I(K.w_aabb + i, e);
        // internal slots are oldest-first; the record is world-list order (newest first)
        for (int r = 0; r < nc; ++r) {
            int k = nc - 1 - r;
            uint32_t* cwp = o + L.off_contacts + MRP_CONTACT_WORDS * r;
            uint32_t m = I(K.w_con + MRP_CONTACT_WORDS * k, e);
            int pc = (m >> 18) & 3;
            cwp[0] = m & 0x000fffffu;
            if (pc == 0) continue;  // stale manifold words are not part of the state
            uint32_t keys = I(K.w_con + MRP_CONTACT_WORDS * k + 1, e);
            cwp[1] = pc == 1 ? (keys & 0xffffu) : keys;
            for (int j = 2; j < 6 + 4 * pc; ++j) cwp[j] = I(K.w_con + MRP_CONTACT_WORDS * k + j, e);
        }
    }
    free(buf);
    return 0;
}

int MRP_API(mrp_set_state)(mrp_handle* h, int32_t env_begin, int32_t env_count, const uint32_t* words) {
    FWD(mrp_set_state_wide(h->wide, env_begin, env_count, words))
    if (!h || !words) return fail(-1, "mrp_set_state: null argument");
    if (env_begin < 0 || env_count < 0 || (int64_t)env_begin + env_count > h->K.N) return fail(-2, "mrp_set_state: bad env range");
    const SimConst& K = h->K;
    const mrp_layout& L = h->L;
    const size_t C = (size_t)env_count;
    uint32_t* buf = (uint32_t*)calloc(C * K.w_total, sizeof(uint32_t));
    if (!buf) return fail(-5, "mrp_set_state: out of host memory");
    auto I = [&](int w, size_t e) -> uint32_t& { return buf[(size_t)w * C + e]; };
    for (size_t e = 0; e < C; ++e) {
        const uint32_t* o = words + e * L.state_words;
        I(W_ELAPSED, e) = o[0]; I(W_EPISODE, e) = o[1]; I(W_INPLACE, e) = o[2];
        int nc = (int)o[3];
        if (nc > K.maxc) { free(buf); return fail(-2, "mrp_set_state: n_contacts exceeds capacity"); }
        I(W_NC, e) = (uint32_t)nc;
        uint32_t gc = 0;
        for (int i = 0; i < K.n; ++i) gc |= (o[L.off_goal_contact + i] ? 1u : 0u) << i;
        I(W_GOALC, e) = gc;
        for (int b = 0; b < K.nb; ++b)
            for (int f = 0; f < 6; ++f) I(K.w_body + kBodyWords * b + f, e) = o[L.off_bodies + 6 * b + f];
        for (int i = 0; i < 2 * (K.n + 1); ++i) I(W_DIST + i, e) = o[L.off_dists + i];
        for (int i = 0; i < 4; ++i) I(W_GOAL + i, e) = o[L.off_goal + i];
        I(W_EPRET, e) = o[L.off_episode_acc]; I(W_EPRET + 1, e) = o[L.off_episode_acc + 1];
        I(W_EPLEN, e) = o[L.off_episode_acc + 2];
        for (int i = 0; i < 4 * K.ndynfix; ++i) I(K.w_aabb + i, e) = o[L.off_aabb + i];
        for (int r = 0; r < nc; ++r) {
            int k = nc - 1 - r;
            const uint32_t* cwp = o + L.off_contacts + MRP_CONTACT_WORDS * r;
            uint32_t m = cwp[0] & 0x000fffffu;
            int fa = m & 0xff, fb = (m >> 8) & 0xff;
            if (fa >= K.nfix || fb >= K.nfix) { free(buf); return fail(-2, "mrp_set_state: bad fixture index"); }
            // body ids of the two fixtures (fixture order: the blocks' fixtures, the robots', the walls)
            auto body_of = [&](int f) { return (int)h->ctab_host[CT_FIXBODY + f]; };
            m |= ((uint32_t)body_of(fa) << 20) | ((uint32_t)body_of(fb) << 24);
            I(K.w_con + MRP_CONTACT_WORDS * k, e) = m;
            for (int j = 1; j < MRP_CONTACT_WORDS; ++j) I(K.w_con + MRP_CONTACT_WORDS * k + j, e) = cwp[j];
        }
    }
    int rc = push_internal(h, env_begin, env_count, buf);
    free(buf);
    if (rc) return fail(-9, "mrp_set_state: copy failed: %s", dev_err());
    return 0;
}

int MRP_API(mrp_enable_terminal_info)(mrp_handle* h, mrp_terminal_buffers* out) {
    FWD(mrp_enable_terminal_info_wide(h->wide, out))
    if (!h || !out) return fail(-1, "mrp_enable_terminal_info: null argument");
    SimConst& K = h->K;
    if (!K.term_obs) {
#ifndef MRP_HOST_EMU
        cudaSetDevice(h->device);
#endif
        const size_t N = (size_t)K.N;
        int rc = DEV_ALLOC(K.term_obs, sizeof(float) * N * K.obs_dim);
        rc |= DEV_ALLOC(K.term_ret, sizeof(float) * N);
        rc |= DEV_ALLOC(K.term_len, sizeof(int32_t) * N);
        if (rc) {
            DEV_FREE(K.term_obs); DEV_FREE(K.term_ret); DEV_FREE(K.term_len);
            K.term_obs = nullptr; K.term_ret = nullptr; K.term_len = nullptr;
            return fail(-7, "mrp_enable_terminal_info: device allocation failed: %s", dev_err());
        }
    }
    out->terminal_obs_dev = K.term_obs;
    out->episode_return_dev = K.term_ret;
    out->episode_length_dev = K.term_len;
    return 0;
}

int MRP_API(mrp_enable_curriculum)(mrp_handle* h, double** scaled_epsilon_dev, double** decay_pow_dev) {
    FWD(mrp_enable_curriculum_wide(h->wide, scaled_epsilon_dev, decay_pow_dev))
    if (!h || !scaled_epsilon_dev || !decay_pow_dev) return fail(-1, "mrp_enable_curriculum: null argument");
    SimConst& K = h->K;
    if (!K.eps_env) {
#ifndef MRP_HOST_EMU
        cudaSetDevice(h->device);
#endif
        const size_t N = (size_t)K.N;
        double *e = nullptr, *d = nullptr;
        if (DEV_ALLOC_RAW(e, sizeof(double) * N) | DEV_ALLOC_RAW(d, sizeof(double) * N)) {
            DEV_FREE(e); DEV_FREE(d);
            return fail(-7, "mrp_enable_curriculum: device allocation failed: %s", dev_err());
        }
        // start from the handle's scalar values
        double* row = (double*)malloc(sizeof(double) * N);
        for (size_t i = 0; i < N; ++i) row[i] = K.rp.scaled_epsilon;
        H2D(e, row, sizeof(double) * N);
        for (size_t i = 0; i < N; ++i) row[i] = K.rp.decay_pow;
        H2D(d, row, sizeof(double) * N);
        free(row);
        K.eps_env = e;
        K.decay_env = d;
    }
    *scaled_epsilon_dev = (double*)K.eps_env;
    *decay_pow_dev = (double*)K.decay_env;
    return 0;
}

int MRP_API(mrp_set_params)(mrp_handle* h, const mrp_params* p) {
    FWD(mrp_set_params_wide(h->wide, p))
    if (!h || !p) return fail(-1, "mrp_set_params: null argument");
    h->K.rp = *p;
    // a spare episode's hidden step evaluates "in place" with the epsilon of the time: recompute them with the new parameters
    if (h->K.spare_ok) {
#ifndef MRP_HOST_EMU
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
#endif
        DEV_ZERO(h->K.spare_ok, (size_t)h->K.N);
    }
    return 0;
}
int MRP_API(mrp_get_params)(mrp_handle* h, mrp_params* p) {
    FWD(mrp_get_params_wide(h->wide, p))
    if (!h || !p) return fail(-1, "mrp_get_params: null argument");
    *p = h->K.rp;
    return 0;
}

int MRP_API(mrp_get_stats)(mrp_handle* h, double* out_host, int32_t reset_after) {
    FWD(mrp_get_stats_wide(h->wide, out_host, reset_after))
    if (!h || !out_host) return fail(-1, "mrp_get_stats: null argument");
#ifndef MRP_HOST_EMU
    cudaSetDevice(h->device);
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(-9, "mrp_get_stats: %s", dev_err());
#endif
    if (D2H(out_host, h->K.stats, sizeof(double) * MRP_N_STATS)) return fail(-9, "mrp_get_stats: D2H failed: %s", dev_err());
    out_host[MRP_STAT_ENV_STEPS] = (double)h->steps_done * (double)h->K.N;  // counted on the host: no per-env atomics
    if (reset_after) { DEV_ZERO(h->K.stats, sizeof(double) * MRP_N_STATS); h->steps_done = 0; }
    return 0;
}

#if defined(MRP_TAILPROBE) && !defined(MRP_WIDE) && !defined(MRP_HOST_EMU)
// profiling builds only: reset (what = 0) or read (what = 1) the tail probe.  out: [kernel][0] start, [kernel][1] longest task
// (ns << 24 | info), [kernel][2 ...] end time of every warp (0 = did not run); row length 2 + kTpWarps
int mrp_debug_tailprobe(int what, unsigned long long* out) {
    if (what == 0) {
        static unsigned long long ones[kTpKernels];
        for (int i = 0; i < kTpKernels; ++i) ones[i] = ~0ull;
        cudaMemcpyToSymbol(g_tp_start, ones, sizeof(ones));
        unsigned long long* p = nullptr;
        cudaGetSymbolAddress((void**)&p, g_tp_end);
        cudaMemset(p, 0, sizeof(unsigned long long) * kTpKernels * kTpWarps);
        cudaGetSymbolAddress((void**)&p, g_tp_maxtask);
        cudaMemset(p, 0, sizeof(unsigned long long) * kTpKernels);
        unsigned int zero = 0;
        cudaMemcpyToSymbol(g_tp_evn, &zero, sizeof(zero));
        return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
    }
    cudaDeviceSynchronize();
    static unsigned long long st[kTpKernels], mx[kTpKernels];
    cudaMemcpyFromSymbol(st, g_tp_start, sizeof(st));
    cudaMemcpyFromSymbol(mx, g_tp_maxtask, sizeof(mx));
    for (int k = 0; k < kTpKernels; ++k) {
        out[(size_t)k * (2 + kTpWarps)] = st[k];
        out[(size_t)k * (2 + kTpWarps) + 1] = mx[k];
        cudaMemcpyFromSymbol(out + (size_t)k * (2 + kTpWarps) + 2, g_tp_end, sizeof(unsigned long long) * kTpWarps,
                             sizeof(unsigned long long) * kTpWarps * k);
    }
    return 0;
}
#endif

#if defined(MRP_TAILPROBE) && !defined(MRP_WIDE) && !defined(MRP_HOST_EMU)
// event-pass records of the last step: returns their number, copies up to cap records of 5 words
int mrp_debug_event_records(unsigned long long* out, int cap) {
    cudaDeviceSynchronize();
    unsigned int n = 0;
    cudaMemcpyFromSymbol(&n, g_tp_evn, sizeof(n));
    if (n > 65536u) n = 65536u;
    const int m = (int)n < cap ? (int)n : cap;
    cudaMemcpyFromSymbol(out, g_tp_evrec, sizeof(unsigned long long) * 5 * (size_t)m);
    return (int)n;
}
#endif

#if defined(MRP_CHECK) && !defined(MRP_HOST_EMU)
// checking build only: failed in-kernel assertions by kind (CHK_*) since the library was loaded
int MRP_API(mrp_debug_check_counts)(unsigned int* out8) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out8, g_check_fail, sizeof(unsigned int) * 8) == cudaSuccess ? 0 : -1;
}
#endif

int64_t MRP_API(mrp_launch_count)(mrp_handle* h) {
#ifndef MRP_WIDE
    if (h && h->wide) return mrp_launch_count_wide(h->wide);
#endif
    return h ? h->launches : 0;
}

}  // extern "C"
