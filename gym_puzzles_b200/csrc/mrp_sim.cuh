// mrp_sim.cuh — one lane = one environment: the whole MultiRobotPuzzle step
// (reference mrp00:413-521 / mrp02:444-584 incl. b2World::Step(1/50, 180, 60)) for the
// sm_100a kernels.  A warp advances a tile of 32 envs; every per-env array that is
// indexed dynamically (bodies, fat AABBs) sits in shared memory in a lane-major layout
// (word k of lane t at warp_block[k*32 + t]) so any divergent index is bank-conflict
// free; the contact cache stays in HBM in a coalesced [word][env] layout.
//
// Box2D semantics followed: SURVEY.md Appendix A (Collide / Solve / SolveTOI ordering,
// contact list = reverse creation order sorted by proxy-id pair, DFS islands over
// newest-first edges, warm starting, block solver, position solver, TOI sub-steps).
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include "../../include/mrp_b200.h"
#include "mrp_collide.cuh"

namespace mrp {

// -DMRP_CHECK (checking build, profiles/check_run.py under gpurun): in-kernel bounds assertions on everything a lane indexes
// dynamically — body field vs. the kernel's shared-memory layout, fixture and contact slots, pool records, queue slots.
// compute-sanitizer is not available on this pool; the host build of the kernel source has the same layout check.
#ifdef MRP_CHECK
#if defined(__CUDACC__)
__device__ unsigned int g_check_fail[8];
MRP_HD void check_fail(int code) {
#if defined(__CUDA_ARCH__)
    atomicAdd(&g_check_fail[code], 1u);
#else
    (void)code;
#endif
}
#define MRP_ASSERT(cond, code) do { if (!(cond)) check_fail(code); } while (0)
#else
#define MRP_ASSERT(cond, code) do { if (!(cond)) { fprintf(stderr, "MRP_CHECK %d failed\n", code); abort(); } } while (0)
#endif
#else
#define MRP_ASSERT(cond, code) do { } while (0)
#endif
enum { CHK_BODY_FIELD = 0, CHK_FIXTURE = 1, CHK_CONTACT_SLOT = 2, CHK_RECORD = 3, CHK_QUEUE = 4, CHK_NC = 5 };

#if defined(__CUDA_ARCH__)
#define MRP_SS 32   // shared-memory lane stride: every warp owns a block of words_per_lane x 32 floats (word k of lane t at k * 32 + t)
#else
#define MRP_SS 1
#endif
constexpr int kBlock = 128;
// Contact slots per env.  The library is compiled for two capacities (see __graft_entry__.build()): 32 — every
// registered variant (measured maximum 21) — and MRP_MAXC = 192 for MultiRobotPuzzle2(num_agents > 2), whose three-fixture
// robots reach >100 simultaneous fat-AABB pairs (188 possible with 5 robots).  Sets of contacts are CMask values.
#ifndef MRP_MAXC
#define MRP_MAXC 32
#endif
constexpr int kMaxC = MRP_MAXC;  // mrp_layout.max_contacts <= kMaxC
// fixtures on dynamic bodies: v0 2 + 8 robots, v2 2 + 3 * 2 in the default build; 2 + 3 * 8 in the wide one
constexpr int kMaxDynFix = MRP_MAXC == 32 ? 12 : 28;
MRP_HD int clz_u32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return __builtin_clz(x);
#endif
}
MRP_HD int ctz_u32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
MRP_HD int popc_u32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
#if MRP_MAXC == 32
typedef uint32_t CMask;
MRP_HD int cm_count(const CMask& m) { return popc_u32(m); }
MRP_HD CMask cm_none() { return 0u; }
MRP_HD CMask cm_all() { return 0xffffffffu; }
MRP_HD void cm_set(CMask& m, int k) { m |= 1u << k; }
MRP_HD void cm_clr(CMask& m, int k) { m &= ~(1u << k); }
MRP_HD bool cm_test(const CMask& m, int k) { return ((m >> k) & 1u) != 0u; }
MRP_HD bool cm_any(const CMask& m) { return m != 0u; }
MRP_HD bool cm_single(const CMask& m) { return m != 0u && (m & (m - 1u)) == 0u; }
MRP_HD int cm_first(const CMask& m) { return ctz_u32(m); }
#else
static_assert(MRP_MAXC % 32 == 0 && MRP_MAXC <= 224, "contact capacity: a multiple of 32, slot index must fit VC_META's 8 bits");
struct CMask { uint32_t w[kMaxC / 32]; };
MRP_HD CMask cm_none() { CMask m; for (int i = 0; i < kMaxC / 32; ++i) m.w[i] = 0u; return m; }
MRP_HD CMask cm_all() { CMask m; for (int i = 0; i < kMaxC / 32; ++i) m.w[i] = 0xffffffffu; return m; }
MRP_HD void cm_set(CMask& m, int k) { m.w[k >> 5] |= 1u << (k & 31); }
MRP_HD void cm_clr(CMask& m, int k) { m.w[k >> 5] &= ~(1u << (k & 31)); }
MRP_HD bool cm_test(const CMask& m, int k) { return ((m.w[k >> 5] >> (k & 31)) & 1u) != 0u; }
MRP_HD bool cm_any(const CMask& m) { uint32_t a = 0u; for (int i = 0; i < kMaxC / 32; ++i) a |= m.w[i]; return a != 0u; }
MRP_HD bool cm_single(const CMask& m) {
    int bits = 0;
    for (int i = 0; i < kMaxC / 32; ++i) for (uint32_t x = m.w[i]; x; x &= x - 1u) ++bits;
    return bits == 1;
}
MRP_HD int cm_first(const CMask& m) {
    for (int i = 0; i < kMaxC / 32; ++i) if (m.w[i]) return 32 * i + ctz_u32(m.w[i]);
    return -1;
}
MRP_HD int cm_count(const CMask& m) { int n = 0; for (int i = 0; i < kMaxC / 32; ++i) n += popc_u32(m.w[i]); return n; }
#endif

// ---- per-CTA constant table (floats; small ints stored as floats) -----------------
constexpr int CT_FIXBODY = 0;     // [32] body index of fixture
constexpr int CT_FIXSHAPE = 32;   // [32] shape index of fixture
constexpr int CT_FIXFRIC = 64;    // [32] friction of fixture
constexpr int CT_WALLFAT = 96;    // [4][4] fat AABB of wall fixtures
constexpr int CT_WALLPOS = 112;   // [4][2] wall body positions
constexpr int kMaxShapes = 10;    // stem, bar, octagon, two wheels, two wall boxes; square variant: + L small, L tall, I
constexpr int CT_SHAPES = 120;    // [kMaxShapes][33]
constexpr int CT_SHAPEX = CT_SHAPES + kMaxShapes * kShapeWords;  // [kMaxShapes][6] bounding data per shape: centre x,y, half extents hx,hy, radius, is_box
constexpr int CT_WALLBOX = CT_SHAPEX + kMaxShapes * 6;          // [4][4] wall polygons in world space (lo.x, lo.y, hi.x, hi.y), no radius
constexpr int CT_BODYP = CT_WALLBOX + 16;                        // [20][4] per body (dynamic bodies, then the four walls: zeros): invMass, invI, local centre x, y
constexpr int CT_WORDS = CT_BODYP + 80;

// ---- internal per-env state words in HBM ---------------------------------------------
// Tiled [tile][word][lane]: envs are grouped in tiles of kTile = 32 (one warp of the per-env kernels); inside a tile
// word w of lane l sits at w * 32 + l.  A warp's access to one word is one 128-byte line (as with a plain [word][env]
// array), but the whole state of a tile is ONE contiguous block of w_total * 128 bytes: consecutive words are neighbouring
// lines of the same DRAM page / TLB entry instead of N * 4 bytes apart, and the word stride is a compile-time constant
// (offsets fold into the load / store instructions).
constexpr int W_ELAPSED = 0, W_EPISODE = 1, W_INPLACE = 2, W_NC = 3, W_GOALC = 4, W_EPLEN = 5;
constexpr int W_EPRET = 6;   // f64
constexpr int W_GOAL = 8;    // 2 x f64
constexpr int W_HINT = 12;   // solver operations this env needed on its previous step (task ordering hint)
constexpr int W_DIST = 13;   // (n+1) x f64: agent_dist[0..n-1], block_dist
// then bodies[nb][11] = {cx, cy, a, vx, vy, w, q.s, q.c, c0x, c0y, a0} (c0/a0: pose at the start of the step, written by
// k_pre for k_post's SynchronizeFixtures); fat[ndynfix][4]; contacts[maxc][14]
constexpr int kBodyWords = 11;
constexpr int kTile = 32;          // envs per state tile
constexpr int kTileShift = 5;

struct SimConst {
    // variant
    int32_t variant, v2, n, nb, nfix, ndynfix, per_agent, maxc, obs_dim, act_dim, max_steps, auto_reset;
    // blocks: bodies 0 .. nblk-1 (1; 3 = T, L, I in the square variant), robots are bodies nblk .. nb-1; the blocks' fixtures come
    // first (nbf of them: 2, or 2 + 2 + 1), then per_agent per robot, then the four walls
    int32_t nblk, nbf;
    int32_t w_body, w_aabb, w_con, w_total;  // word offsets of the internal state
    int32_t smem_words;                      // per-lane shared-memory words
    int32_t stage_rows;                      // 1 (device default): action rows in / observation rows out move through shared memory as coalesced runs (k_pre, k_post)
    int32_t big_split;                       // 1: islands of class 3 (more than two contacts) are solved by k_solve_big, not by k_solve_vel / k_solve_pos
    // body classes
    float blk_mass, blk_invMass, blk_invI, blk_lcx, blk_lcy;
    float ag_mass, ag_invMass, ag_invI, ag_lcx, ag_lcy, ag_inertia;
    float lin_k, ang_k;  // 1/(1+h*damping)
    float h;
    float blkv[19][2]; // observation vertex list (bar verts, then stem verts; square variant: T 8, L 7, I 4 — blkv_off)
    int32_t blkv_off[4];
    // square variant: mass data of blocks 1, 2 (block 0 is blk_*), target COM of block k relative to the goal centre (obs units)
    // and target angle
    float blkx_invMass[3], blkx_invI[3], blkx_lcx[3], blkx_lcy[3];
    double sq_target[3][3];
    // env constants (float64, as the reference's Python arithmetic)
    double SCALE, W, H, ratio, SPEED;
    double goal_x0, goal_y0;
    mrp_params rp;
    uint64_t seed, env_id_base;
    // memory
    uint32_t* S;
    int64_t N;       // envs of the handle
    int64_t env0;    // first env of the chunk this launch works on
    int32_t nloc;    // envs in the chunk (chunks are pipelined on separate streams; queues and counters are per chunk)
    int32_t pad1;
    const float* ctab;
    const float* act;
    float* obs;
    float* rew;
    uint8_t* done;
    uint8_t* trunc;
    double* stats;
    int32_t* reset_list;
    const uint8_t* reset_mask;
    // spare episodes (large batches): the state and first observation of every env's NEXT episode, computed ahead of time beside
    // the step's kernels (a respawn depends on seed, env id and episode number only), so that the auto-reset of a finished env is a
    // copy instead of a whole fused step on one lane at the end of the step
    uint32_t* S2;          // spare states, same tiled layout as S (NULL: no spares, auto-reset runs reset_lane)
    float* obs2;           // [N][obs_dim] observation the reset returns
    uint8_t* spare_ok;     // [N] 1: the spare of this env is valid
    int32_t* refill_list;  // envs whose spare has to be (re)computed, [min(refill_cnt[0], refill_cap)] entries
    int32_t* refill_cnt;
    int32_t refill_cap;    // spares computed per step at most (the rest waits for the next step)
    // the refill pass runs the phase pipeline over a LIST of envs (idx_list, *idx_count entries, at most nloc) with S = S2,
    // obs = obs2 and `hidden` set: the step after a respawn — action drawn from the reset stream, no reward / done / TimeLimit
    // accounting, the spare marked valid at the end
    const int32_t* idx_list;
    const int32_t* idx_count;
    int32_t hidden;
    int32_t pad2;
    // phase pipeline (DESIGN.md "kernels"): solver tasks produced by k_pre, consumed by k_solve_vel / k_solve_pos
    float* pool;          // constraint records, VC_WORDS floats each, allocated per task with one atomic
    int32_t* cnt;         // [CNT_*] counters, zeroed at the start of every step
    int32_t* task_env;    // [N] env of task i
    int32_t* task_T;      // [N] number of constraints of task i
    int32_t* task_off;    // [..] index of task i's first record in pool
    int32_t* toi_list;    // [N] envs whose TOI scan found an event (handled by k_post_events)
    uint32_t* narrow_list; // [N * maxc] contacts that need SAT + clipping this step: env * kMaxC + slot
    int32_t* post_list;    // [N] envs without solver tasks from slot 0 up, envs with tasks from the last slot down (k_pre):
                           // k_post of the former runs beside the solver kernels
    // optional per-env curriculum vectors (NULL: the scalar mrp_params apply): update_goal / update_params per env
    const double* eps_env;       // [N] scaled_epsilon   (mrp02:232-233)
    const double* decay_env;     // [N] decay**(-timestep) (mrp02:227-230)
    // optional terminal records of envs that finished inside this step, written before the auto-reset overwrites them
    float* term_obs;             // [N][obs_dim] last observation of the finished episode (SB3 "terminal_observation")
    float* term_ret;             // [N] return of the finished episode (Monitor info["episode"]["r"])
    int32_t* term_len;           // [N] its length                   (info["episode"]["l"])
};
// solver tasks predicted heavy (by W_HINT) are queued from slot 0 upwards, the others from slot N-1 downwards;
// consumers take the heavy ones first so that long solves start early and short ones fill the tail
// Solver tasks come in three classes so that the lanes of a warp run the same instruction stream:
//   0: one contact, 1-point manifold   1: one contact, 2-point manifold (block solver)   2: two contacts   3: more.
// Class c owns slots [c*cap, (c+1)*cap), cap = N * nb.
constexpr int kTaskClasses = 4;
enum { CNT_RESET = 0, CNT_POOL = 1, CNT_TOI = 2, CNT_NARROW = 3, CNT_HEAD_P = 4, CNT_TASKS = 8 /*[4] heavy*/, CNT_TASKS_LIGHT = 12 /*[4]*/,
       CNT_HEAD_V = 16 /*[4]*/, CNT_FREE = 20 /* envs without solver tasks */, CNT_BUSY = 21 /* envs with tasks */,
       CNT_TOI_F = 22 /* TOI-event queue of the task-free group (filled from the end of toi_list) */, CNT_N = 23 };
// transient meta bits used between k_broad, k_narrow and k_pre (cleared again by k_pre)
constexpr uint32_t kMetaWas = 1u << 29, kMetaDead = 1u << 30;
constexpr uint32_t kHeavyHint = 120;
// address of word 0 of env `env` (local index within the handle); word w is at [w * kTile]
MRP_HD uint32_t* env_words(const SimConst& K, int64_t env) {
    return K.S + (env >> kTileShift) * ((int64_t)K.w_total << kTileShift) + (env & (kTile - 1));
}
MRP_HD int task_count(const SimConst& K, int cls) { return K.cnt[CNT_TASKS + cls] + K.cnt[CNT_TASKS_LIGHT + cls]; }
MRP_HD int task_slot(const SimConst& K, int cls, int i) {  // i-th task of a class in consumption order -> slot
    const int cap = K.nloc * K.nb;
    const int nh = K.cnt[CNT_TASKS + cls];
    return cls * cap + (i < nh ? i : cap - 1 - (i - nh));
}
MRP_HD int task_slot_any(const SimConst& K, int i) {  // i-th task over all classes, heaviest class first (class 3 left out under big_split)
    for (int cls = kTaskClasses - 1 - (K.big_split ? 1 : 0); cls > 0; --cls) {
        const int n = task_count(K, cls);
        if (i < n) return task_slot(K, cls, i);
        i -= n;
    }
    return task_slot(K, 0, i);
}
MRP_HD int task_count_all(const SimConst& K) {
    int n = 0;
    for (int cls = 0; cls < kTaskClasses - (K.big_split ? 1 : 0); ++cls) n += task_count(K, cls);
    return n;
}

MRP_HD int atomic_add_i32(int32_t* p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    int o = *p;
    *p = o + v;
    return o;
#endif
}

// ---- solver constraint record (per touching contact, in island order) --------------
constexpr int VC_META = 0;   // bA | bB<<4 | vpc<<8 | ppc<<10 | type<<12 | island<<16 | slot<<24
constexpr int VC_FRIC = 1, VC_NX = 2, VC_NY = 3;
constexpr int VC_PT = 4;     // [2][8]: rAx rAy rBx rBy nMass tMass nI tI
constexpr int VC_K11 = 20, VC_K12 = 21, VC_K22 = 22, VC_M11 = 23, VC_M12 = 24, VC_M22 = 25;
constexpr int VC_LNX = 26, VC_LNY = 27, VC_LPX = 28, VC_LPY = 29, VC_LP0X = 30, VC_LP0Y = 31, VC_LP1X = 32, VC_LP1Y = 33;
constexpr int VC_MA = 34, VC_IA = 35, VC_MB = 36, VC_IB = 37;  // invMassA invIA invMassB invIB
constexpr int VC_WORDS = 38;

constexpr int kDynFields = 17;  // per-lane shared-memory words of a dynamic body in the per-env kernels

struct Sim {
    const SimConst& K;
    float* sm;          // this lane's shared-memory column
    const float* ct;    // CTA constant table (shared memory)
    uint32_t* G;        // word 0 of the env's state (re-pointed per task in the solver kernels)
    int32_t env_i;      // env index within the handle
    uint64_t gid;
    // per-lane scratch (local memory where dynamically indexed)
    uint32_t meta[kMaxC];  // fa | fb<<8 | touching<<16 | type<<17 | pc<<18 | bA<<20 | bB<<24
    float* vcp;            // solver constraint records: a lane-local array (fused path) or a slot of the HBM task pool
    int fdyn;              // shared-memory words per dynamic body
    int qoff;              // first of the three rotation-cache words inside a body slot
    int fa_off;            // first shared-memory word of the fat AABBs
    int c0f;               // first of the four words {c0x, c0y, a0, alpha0} inside a body slot (-1: absent)
    int wall_off;          // first shared-memory word of the wall slots (-1: absent, walls are read from the constant table)
    uint8_t order[kMaxC];
    float toi[kMaxC];
    uint8_t toiCount[kMaxC];
    float wallAlpha0[4];
    float alpha_none;      // alpha0 of every dynamic body in the k_post layout (no TOI event is processed there)
    float swept[kMaxDynFix * 4];   // swept tight AABB (incl. polygon radius) of every dynamic fixture, from SynchronizeFixtures
    int nc;
    uint32_t goalc;
    uint32_t overflow;
    // observation output (obs_put): straight into the env's row, or — k_post on the device — staged through the warp's shared
    // memory and written out by the whole warp as coalesced float4 runs (obs_flush)
    float* orow;           // this env's row of the [N][obs_dim] buffer
    int opos;              // values written to the row so far
    float* ost;            // this lane's column of the warp's staging block (nullptr: direct row writes)
    int owin, ocnt;        // staging window length, values of the current window staged so far
    unsigned omask;        // lanes of the warp that produce a row in this pass
    bool stored;           // store() already ran (the staging block lies over the fat-AABB words, dead after store())
    // workload counters of this lane (MRP_STAT_M1 / M2 / POS_POINTS / TOI_CALLS): summed per warp by the kernels
    uint32_t stat_m1, stat_m2, stat_pos_pts, stat_toi;
#ifdef MRP_TAILPROBE
    long long tp_toi_clk = 0, tp_evt_clk = 0;   // cycles inside time_of_impact / toi_event of this lane
    int tp_evt_n = 0;
    static MRP_HD long long tp_clock() {
#if defined(__CUDA_ARCH__)
        return clock64();
#else
        return 0;
#endif
    }
#endif

    // layouts (words per dynamic body):
    //   17 full: pose/vel 0-5, q 6-7, p 8-9, cache 10-12, c0/a0/alpha0 13-16, walls, fat AABBs   (fused step, reset, k_post_events)
    //   11 k_post: 0-7, c0/a0 8-10, fat AABBs; no p (recomputed from c, q), no alpha0 (0: no TOI event is processed
    //      in this layout), no cache, no wall slots
    //    8 k_pre: 0-7, walls (origin recomputed, rotation = q)   10 k_broad: 0-9, fat AABBs
    //    9 position solver: 0-5, cache 6-8, walls               6 velocity solver: 0-5, walls
    MRP_HD Sim(const SimConst& k, float* sm_, const float* ct_, int64_t env, float* vc_ = nullptr, int fdyn_ = kDynFields)
        : K(k), sm(sm_), ct(ct_), G(env_words(k, env)), env_i((int32_t)env), gid(k.env_id_base + (uint64_t)env), vcp(vc_), fdyn(fdyn_),
          qoff(fdyn_ == 17 || fdyn_ == 13 ? 10 : (fdyn_ == 9 ? 6 : -1)),
          fa_off(fdyn_ == 17 ? k.nb * 17 + 24 : (fdyn_ == 11 ? k.nb * 11 : (fdyn_ == 10 ? k.nb * 10 : -1))),
          c0f(fdyn_ == 17 ? 13 : (fdyn_ == 11 ? 8 : -1)), wall_off(fdyn_ == 11 || fdyn_ == 10 ? -1 : k.nb * fdyn_), nc(0), goalc(0),
          overflow(0), orow(nullptr), opos(0), ost(nullptr), owin(0), ocnt(0), omask(0u), stored(false), stat_m1(0), stat_m2(0),
          stat_pos_pts(0), stat_toi(0) {}

    // ------------------------------------------------------------ memory helpers
    MRP_HD uint32_t& g(int w) { return G[w << kTileShift]; }
    MRP_HD void set_env(int64_t env) { G = env_words(K, env); env_i = (int32_t)env; }
    MRP_HD float gf(int w) { return __uint_as_float_(g(w)); }
    MRP_HD void gsf(int w, float v) { g(w) = __float_as_uint_(v); }
    MRP_HD double gd(int w) {
        uint64_t lo = g(w), hi = g(w + 1);
        return __ull_as_double_(lo | (hi << 32));
    }
    MRP_HD void gsd(int w, double v) {
        uint64_t u = __double_as_ull_(v);
        g(w) = (uint32_t)u;
        g(w + 1) = (uint32_t)(u >> 32);
    }
    static MRP_HD float __uint_as_float_(uint32_t u) { union { uint32_t u; float f; } c; c.u = u; return c.f; }
    static MRP_HD uint32_t __float_as_uint_(float f) { union { uint32_t u; float f; } c; c.f = f; return c.u; }
    static MRP_HD double __ull_as_double_(uint64_t u) { union { uint64_t u; double f; } c; c.u = u; return c.f; }
    static MRP_HD uint64_t __double_as_ull_(double f) { union { uint64_t u; double f; } c; c.f = f; return c.u; }
    MRP_HD int cw(int k, int j) const {
        MRP_ASSERT(k >= 0 && k < K.maxc && j >= 0 && j < MRP_CONTACT_WORDS, CHK_CONTACT_SLOT);
        return K.w_con + k * MRP_CONTACT_WORDS + j;
    }

    // body fields: 0 cx 1 cy 2 a 3 vx 4 vy 5 w (all bodies incl. walls) | 6 qs 7 qc 8 px 9 py | 10 cache.s 11 cache.c
    // 12 cache angle (last Rot evaluated for this body and the angle it belongs to) | 13 c0x 14 c0y 15 a0 16 alpha0.
    // Kernels that do not need the tail allocate fewer words per body (fdyn): k_broad 10, k_pre 8, k_post 11, event / reset paths 17;
    // the solver kernels use 6 (velocities) and 9 (pose + cache at qoff = 6).
    // body-major, lane-strided: one base computation per body, then compile-time field offsets
    MRP_HD float* bp(int b) { return sm + (b < K.nb ? b * fdyn : wall_off + (b - K.nb) * 6) * MRP_SS; }
    MRP_HD V2 wall_pos(int b) const { return mk(ct[CT_WALLPOS + 2 * (b - K.nb)], ct[CT_WALLPOS + 2 * (b - K.nb) + 1]); }
#ifdef MRP_HOST_EMU
    // the host build checks that a kernel never touches a body word its shared-memory layout does not have
    float& B(int b, int f) {
        if (b < K.nb ? f >= fdyn : (f >= 6 || wall_off < 0)) { fprintf(stderr, "smem layout violation: body %d field %d fdyn %d\n", b, f, fdyn); abort(); }
        return bp(b)[f * MRP_SS];
    }
    float& BX(int b, int f) { return B(b, f); }
#else
    MRP_HD float& B(int b, int f) {
        MRP_ASSERT(b >= 0 && b < K.nb + 4 && (b < K.nb ? f < fdyn : (f < 6 && wall_off >= 0)), CHK_BODY_FIELD);
        return bp(b)[f * MRP_SS];
    }
    MRP_HD float& BX(int b, int f) { return B(b, f); }
#endif
    MRP_HD float& FA(int fx, int j) {
        MRP_ASSERT(fa_off >= 0 && fx >= 0 && fx < K.ndynfix && j >= 0 && j < 4, CHK_FIXTURE);
        return sm[(fa_off + fx * 4 + j) * MRP_SS];
    }

    // b2Rot::Set(angle of body b) through a one-entry cache: sin/cos are pure functions of the float angle, so
    // reusing the last evaluation when the angle is bit-identical cannot change results (robots with invI = 0
    // never rotate inside the position solver; the block only when an impulse acts on it)
    MRP_HD void set_rot_cache(int b, Rot q, float angle) {
        if (qoff < 0) return;
        float* const c = bp(b) + qoff * MRP_SS;
        c[0] = q.s; c[MRP_SS] = q.c; c[2 * MRP_SS] = angle;
    }
    template <bool INL = false>
    MRP_HD Rot body_rot(int b, float angle) {
        Rot q;
        if (b >= K.nb) { q.s = 0.0f; q.c = 1.0f; return q; }
        if (fdyn == 8) return Rot{BX(b, 6), BX(b, 7)};   // k_pre: angles are the loaded ones throughout, q belongs to them
        if (qoff < 0) return INL ? rot_set_inline(angle) : rot_set(angle);
        float* const c = bp(b) + qoff * MRP_SS;
        if (c[2 * MRP_SS] == angle) { q.s = c[0]; q.c = c[MRP_SS]; return q; }
        q = INL ? rot_set_inline(angle) : rot_set(angle);
        c[0] = q.s; c[MRP_SS] = q.c; c[2 * MRP_SS] = angle;
        return q;
    }

    MRP_HD bool is_dyn(int b) const { return b < K.nb; }
    // one table read each (CT_BODYP, filled by fill_variant from blk_* / blkx_* / ag_*) instead of a chain of selects over kernel parameters
    MRP_HD float invMass(int b) const { return ct[CT_BODYP + 4 * b]; }
    MRP_HD float invI(int b) const { return ct[CT_BODYP + 4 * b + 1]; }
    MRP_HD V2 localCenter(int b) const { return mk(ct[CT_BODYP + 4 * b + 2], ct[CT_BODYP + 4 * b + 3]); }
    // fixtures [f0, f1) of dynamic body b (creation order: the blocks' fixtures, then the robots')
    MRP_HD void fix_range(int b, int& f0, int& f1) const {
        if (b >= K.nblk) { f0 = K.nbf + K.per_agent * (b - K.nblk); f1 = f0 + K.per_agent; }
        else if (K.nblk == 1) { f0 = 0; f1 = 2; }
        else { f0 = 2 * b; f1 = b == 2 ? 5 : f0 + 2; }   // T 0-1, L 2-3, I 4
    }
    // body of the current goal block: the block queue is T, L, I, so its head is the number of blocks already in place
    MRP_HD int goal_block() {
        if (K.nblk == 1) return 0;
        const int placed = (int)g(W_INPLACE);
        return placed < K.nblk - 1 ? placed : K.nblk - 1;
    }
    MRP_HD Xf body_xf(int b) {
        Xf x;
        if (b < K.nb) {
            if (fdyn == 8 && K.hidden) {   // first step after a spawn: the body origin as spawned (see load())
                x.p = mk(gf(K.w_body + kBodyWords * b + 8), gf(K.w_body + kBodyWords * b + 9));
            } else if (fdyn == 11 || fdyn == 8) {  // p = c - R(q) * localCenter: the very expression sync_transform stores
                const V2 r = rmul(Rot{BX(b, 6), BX(b, 7)}, localCenter(b));
                x.p = mk(B(b, 0) - r.x, B(b, 1) - r.y);
            } else {
                x.p = mk(BX(b, 8), BX(b, 9));
            }
            x.q.s = BX(b, 6);
            x.q.c = BX(b, 7);
        } else {
            x.p = wall_pos(b);
            x.q.s = 0.0f;
            x.q.c = 1.0f;
        }
        return x;
    }
    MRP_HD void sync_transform(int b) {  // b2Body::SynchronizeTransform
        Rot q = body_rot(b, B(b, 2));
        V2 r = rmul(q, localCenter(b));
        BX(b, 6) = q.s;
        BX(b, 7) = q.c;
        if (fdyn != 11 && fdyn != 8) {
            BX(b, 8) = B(b, 0) - r.x;
            BX(b, 9) = B(b, 1) - r.y;
        }
    }
    MRP_HD int fix_body(int f) const { return (int)ct[CT_FIXBODY + f]; }
    MRP_HD const float* fix_shape(int f) const { return ct + CT_SHAPES + kShapeWords * (int)ct[CT_FIXSHAPE + f]; }
    MRP_HD Box fat(int f) {
        Box b;
        if (f < K.ndynfix) {
            b.lx = FA(f, 0); b.ly = FA(f, 1); b.hx = FA(f, 2); b.hy = FA(f, 3);
        } else {
            const float* w = ct + CT_WALLFAT + 4 * (f - K.ndynfix);
            b.lx = w[0]; b.ly = w[1]; b.hx = w[2]; b.hy = w[3];
        }
        return b;
    }
    MRP_HD float& alpha0(int b) { return b < K.nb ? (fdyn == 11 ? alpha_none : BX(b, c0f + 3)) : wallAlpha0[b - K.nb]; }

    // ------------------------------------------------------------ state load / store
    // Device: the body and fat-AABB words go from the state straight into the lane's shared-memory column with cp.async
    // (LDGSTS, 4 bytes each): no register is tied up waiting, so ALL of them are in flight at once and the load phase costs
    // one memory round trip (plus one for the contact heads, whose count has to arrive first) instead of one per body /
    // fixture pair.  Host build: plain copies.
#if defined(__CUDA_ARCH__)
    static __device__ __forceinline__ void cp_word(float* dst_smem, const uint32_t* src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
    }
    static __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
#else
    static void cp_word(float* dst, const uint32_t* src) { union { uint32_t u; float f; } c; c.u = *src; *dst = c.f; }
    static void cp_wait() {}
#endif
    MRP_HD void init_walls() {
        for (int k = 0; wall_off >= 0 && k < 4; ++k) {
            int b = K.nb + k;
            B(b, 0) = ct[CT_WALLPOS + 2 * k];
            B(b, 1) = ct[CT_WALLPOS + 2 * k + 1];
            B(b, 2) = 0.0f; B(b, 3) = 0.0f; B(b, 4) = 0.0f; B(b, 5) = 0.0f;
        }
    }
    // walls = false: the wall slots are left alone (k_pre stages the warp's action rows there; init_walls() follows)
    MRP_HD void load(bool walls = true) {
        // words a kernel's layout has no use for are not fetched: k_broad (10) needs pose and rotation only, k_pre (13)
        // writes the pre-step pose words itself
        const bool want_vel = fdyn != 10, want_c0 = c0f >= 0;
        for (int b = 0; b < K.nb; ++b) {
            const uint32_t* src = &g(K.w_body + kBodyWords * b);
            float* p = bp(b);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < 3 || i > 5 || want_vel) cp_word(p + i * MRP_SS, src + (i << kTileShift));
            if (want_c0) {
#pragma unroll
                for (int i = 0; i < 3; ++i) cp_word(p + (c0f + i) * MRP_SS, src + ((8 + i) << kTileShift));  // pre-step pose
            }
        }
        for (int f = 0; fa_off >= 0 && f < K.ndynfix; ++f) {
            const uint32_t* src = &g(K.w_aabb + 4 * f);
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_word(&FA(f, i), src + (i << kTileShift));
        }
        nc = (int)g(W_NC);
        goalc = g(W_GOALC);
        for (int k = 0; k < nc; k += 4) {
            uint32_t r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) r[i] = k + i < nc ? g(cw(k + i, 0)) : 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) if (k + i < nc) meta[k + i] = r[i];
        }
        cp_wait();
        for (int b = 0; b < K.nb; ++b) {
            float* p = bp(b);
            if (!want_vel) { p[3 * MRP_SS] = 0.0f; p[4 * MRP_SS] = 0.0f; p[5 * MRP_SS] = 0.0f; }
            const Rot q{p[6 * MRP_SS], p[7 * MRP_SS]};
            set_rot_cache(b, q, p[2 * MRP_SS]);
            if (fdyn == 8) {
                // k_pre keeps no origin in shared memory: body_xf() recomputes it (or, for a hidden step, reads the spawn's)
            } else if (K.hidden && fdyn == 10) {
                // first step after a spawn: the body origin is the spawn position itself (b2Body's constructor), kept in the
                // pre-step pose words by spawn_spare_lane until k_pre overwrites them; c - R(q) * localCenter rounds differently
                p[8 * MRP_SS] = gf(K.w_body + kBodyWords * b + 8);
                p[9 * MRP_SS] = gf(K.w_body + kBodyWords * b + 9);
            } else if (fdyn != 11) {
                V2 rc = rmul(q, localCenter(b));
                p[8 * MRP_SS] = p[0] - rc.x;
                p[9 * MRP_SS] = p[MRP_SS] - rc.y;
            }
        }
        if (walls) init_walls();
    }

    // ---- observation rows -----------------------------------------------------------------------------------------
    // A lane produces the values of its env's row one after the other.  Written straight to the row, every store
    // instruction of the warp touches 32 rows (32 sectors for 128 bytes).  In k_post the values are staged instead:
    // lane t puts value c of the current window at st[c * 33 + t] (conflict-free), and when the window is full the warp
    // writes it out together, consecutive lanes to consecutive float4 of a row (obs_flush).  The staging block lies over
    // the warp's fat-AABB words, which are dead once store() has run: no extra shared memory, occupancy unchanged.
    static constexpr int kObsPad = 33;
    MRP_HD void obs_begin(float* row) { orow = row; opos = 0; ocnt = 0; }
    MRP_HD void obs_put(float v) {
#if defined(__CUDA_ARCH__)
        if (ost) {
            ost[ocnt * kObsPad] = v;
            if (++ocnt == owin) obs_flush();
            return;
        }
#endif
        orow[opos++] = v;
    }
    MRP_HD void obs_end() {
#if defined(__CUDA_ARCH__)
        if (ost && ocnt) obs_flush();
#endif
    }
#if defined(__CUDA_ARCH__)
    // window length for a staging block of `words` floats per lane (the 32 env indices sit behind the window); float4 runs
    // need windows that start and end on multiples of four values.  Rows that are not a multiple of 16 bytes (v2 with its two
    // default robots: 39 values) are written directly: their staged variant (coalesced scalar runs) was measured 3 % slower
    // than the direct stores (1.87 -> 1.93 ms per step of 1M envs) — k_post is not bound by store bandwidth
    __device__ static int obs_window(int words, int obs_dim) {
        if (obs_dim & 3) return 0;
        const int w = ((words * 32 - 32) / kObsPad) & ~3;
        return w < obs_dim ? w : obs_dim;
    }
    // called by every lane of the warp that entered the pass (entry), converged or not; fin = this lane produces a row.
    // Runs store() and turns the fat-AABB words of the warp into the staging block.
    __device__ __forceinline__ void obs_stage(unsigned entry, bool fin) {
        omask = __ballot_sync(entry, fin);
        if (!fin) return;
        owin = obs_window(4 * K.ndynfix, K.obs_dim);
        if (owin < 4) return;               // no room (never with the registered variants): direct row writes
        store();
        stored = true;
        __syncwarp(omask);                  // every lane's AABB words are in the state before anyone overwrites them
        const int lane = threadIdx.x & 31;
        float* blk = sm - lane + fa_off * 32;
        ost = blk + lane;
        reinterpret_cast<int32_t*>(blk + owin * kObsPad)[lane] = env_i;
    }
    __device__ __noinline__ void obs_flush() {
        __syncwarp(omask);
        const int lane = threadIdx.x & 31, cnt = ocnt;
        const int rank = __popc(omask & ((1u << lane) - 1u)), nact = __popc(omask);
        const float* blk = ost - lane;
        const int32_t* envs = reinterpret_cast<const int32_t*>(blk + owin * kObsPad);
        const int q = cnt >> 2, total = 32 * q;     // cnt, opos and obs_dim are multiples of four (obs_window)
        for (int idx = rank; idx < total; idx += nact) {
            const int row = idx / q, c = (idx - row * q) * 4;
            if (!((omask >> row) & 1u)) continue;
            const float* v = blk + c * kObsPad + row;
            *reinterpret_cast<float4*>(K.obs + (int64_t)envs[row] * K.obs_dim + opos + c) = make_float4(v[0], v[kObsPad], v[2 * kObsPad], v[3 * kObsPad]);
        }
        __syncwarp(omask);                  // the window is reused, and the rows are visible to the warp (terminal-obs copy)
        opos += cnt;
        ocnt = 0;
    }
#endif
    MRP_HD void store() {
        g(W_NC) = (uint32_t)nc;
        g(W_GOALC) = goalc;
        for (int b = 0; b < K.nb; ++b) {
            int w = K.w_body + kBodyWords * b;
            for (int f = 0; f < 6; ++f) gsf(w + f, B(b, f));
            gsf(w + 6, BX(b, 6));
            gsf(w + 7, BX(b, 7));
        }
        for (int f = 0; f < K.ndynfix; ++f)
            for (int j = 0; j < 4; ++j) gsf(K.w_aabb + 4 * f + j, FA(f, j));
        for (int k = 0; k < nc; ++k) g(cw(k, 0)) = meta[k];
    }

    // ------------------------------------------------------------ contact events (ContactDetector, mrp00:92-111)
    MRP_HD void contact_event(uint32_t m, bool begin) {
        int bA = (m >> 20) & 15, bB = (m >> 24) & 15;
        const int gb = goal_block();
        int ag = bA == gb ? bB : (bB == gb ? bA : -1);
        if (ag >= K.nblk && ag < K.nb) {
            uint32_t bit = 1u << (ag - K.nblk);
            goalc = begin ? (goalc | bit) : (goalc & ~bit);
        }
    }

    // ---- exact culling -------------------------------------------------------------------------------------------
    // b2CollidePolygons returns an empty manifold as soon as one polygon's best face separation exceeds
    // totalRadius (0.02).  A cheap LOWER bound on that separation that already exceeds 0.02 (plus a guard band far
    // above float rounding) therefore proves the manifold empty without running SAT + clipping:
    //  * wall pairs: the wall's face normals are the world axes, so the separation along them is an AABB gap;
    //  * box fixtures (T-block parts, v2 wheels): the box's own face normals against the other shape's bounding circle;
    //  * octagon vs octagon: face normals are 45 degrees apart, so one lies within 22.5 degrees of the centre line.
    MRP_HD const float* shape_x(int f) const { return ct + CT_SHAPEX + 6 * (int)ct[CT_FIXSHAPE + f]; }
    // half extents of fixture f's polygon along the world axes at rotation q (bounding radius for the octagon)
    MRP_HD V2 world_extent(const float* sx, Rot q) const {
        if (sx[5] == 0.0f) return mk(sx[4], sx[4]);
        const float ac = fabsf(q.c), as = fabsf(q.s);
        return mk(ac * sx[2] + as * sx[3], as * sx[2] + ac * sx[3]);
    }
    MRP_HD bool manifold_provably_empty(int fa, int fb, int bA, int bB, Xf xfA, Xf xfB) {
        const float kGuard = 0.0205f;  // totalRadius 0.02 + guard
        const float* sa = shape_x(fa);
        const float* sb = shape_x(fb);
        const V2 cA = xmul(xfA, mk(sa[0], sa[1]));
        if (bB >= K.nb) {  // fb is a wall (walls are the last fixtures)
            const float* w = ct + CT_WALLBOX + 4 * (bB - K.nb);
            const V2 e = world_extent(sa, xfA.q);
            const float gx = fmax2(w[0] - (cA.x + e.x), (cA.x - e.x) - w[2]);
            const float gy = fmax2(w[1] - (cA.y + e.y), (cA.y - e.y) - w[3]);
            return fmax2(gx, gy) > kGuard;
        }
        const V2 cB = xmul(xfB, mk(sb[0], sb[1]));
        if (sa[5] != 0.0f) {  // A is a box: its face normals vs B's bounding circle
            const V2 p = rmulT(xfA.q, cB - cA);
            if (fmax2(fabsf(p.x) - sa[2], fabsf(p.y) - sa[3]) - sb[4] > kGuard) return true;
        }
        if (sb[5] != 0.0f) {
            const V2 p = rmulT(xfB.q, cA - cB);
            if (fmax2(fabsf(p.x) - sb[2], fabsf(p.y) - sb[3]) - sa[4] > kGuard) return true;
        }
        if (sa[5] == 0.0f && sb[5] == 0.0f) {
            const float D = length(cB - cA);
            if (0.9238f * D - sa[4] - sb[4] > kGuard) return true;
        }
        return false;
    }

    // b2Contact::Update for slot k (A.4): narrowphase at the bodies' current transforms, impulse
    // matching by feature id, touching flag, Begin/End events.
    MRP_HDN void update_contact(int k) {
        uint32_t m = meta[k];
        int fa = m & 0xff, fb = (m >> 8) & 0xff;
        int bA = (m >> 20) & 15, bB = (m >> 24) & 15;
        bool was = (m >> 16) & 1;
        int oldpc = (m >> 18) & 3;
        Manifold man;
        const Xf xfA = body_xf(bA), xfB = body_xf(bB);
        if (manifold_provably_empty(fa, fb, bA, bB, xfA, xfB)) { man.pc = 0; man.type = 0; }
        else collide_polygons(&man, fix_shape(fa), xfA, fix_shape(fb), xfB);
        bool touching = man.pc > 0;
        if (touching) {
            uint32_t okeys = oldpc ? g(cw(k, 1)) : 0u;
            float nI[2] = {0.0f, 0.0f}, tI[2] = {0.0f, 0.0f};
            for (int i = 0; i < man.pc; ++i) {
                for (int j = 0; j < oldpc; ++j) {
                    if (((okeys >> (16 * j)) & 0xffffu) == man.key[i]) {
                        nI[i] = gf(cw(k, 8 + 4 * j));
                        tI[i] = gf(cw(k, 9 + 4 * j));
                        break;
                    }
                }
            }
            g(cw(k, 1)) = man.key[0] | ((man.pc > 1 ? man.key[1] : 0u) << 16);
            gsf(cw(k, 2), man.ln.x); gsf(cw(k, 3), man.ln.y);
            gsf(cw(k, 4), man.lp.x); gsf(cw(k, 5), man.lp.y);
            for (int i = 0; i < man.pc; ++i) {
                gsf(cw(k, 6 + 4 * i), man.pt[i].x);
                gsf(cw(k, 7 + 4 * i), man.pt[i].y);
                gsf(cw(k, 8 + 4 * i), nI[i]);
                gsf(cw(k, 9 + 4 * i), tI[i]);
            }
        }
        m = (m & 0xfff0ffffu) | ((touching ? 1u : 0u) << 16) | (((uint32_t)man.type & 1u) << 17) | ((uint32_t)man.pc << 18);
        if (!touching) m &= ~(1u << 17);
        meta[k] = m;
        if (!was && touching) contact_event(m, true);
        if (was && !touching) contact_event(m, false);
    }

    // b2ContactManager::Collide (A.3): world-list order = newest slot first
    // ---- Collide split over three kernels: k_broad classifies every contact (destroyed / provably empty / needs
    // SAT), k_narrow runs SAT + clipping per queued contact with all lanes busy, k_pre replays the Begin/End events in
    // contact-list order and compacts.  The three together are exactly collide() below.
    // returns the set of contacts that need SAT + clipping (the caller queues them for k_narrow)
    MRP_HD CMask broad_phase() {
        CMask need = cm_none();
        for (int k = nc - 1; k >= 0; --k) {
            uint32_t m = meta[k] & 0x0fffffffu;
            if ((m >> 16) & 1) m |= kMetaWas;
            const int fa = m & 0xff, fb = (m >> 8) & 0xff;
            const int bA = (m >> 20) & 15, bB = (m >> 24) & 15;
            if (!overlap(fat(fa), fat(fb))) {
                m |= kMetaDead;
            } else if (manifold_provably_empty(fa, fb, bA, bB, body_xf(bA), body_xf(bB))) {
                m &= 0xfff0ffffu | kMetaWas;  // pointCount 0, not touching
            } else {
                cm_set(need, k);
            }
            g(cw(k, 0)) = m;
        }
        return need;
    }
    MRP_HD void finish_collide() {
        CMask dead = cm_none();
        for (int k = nc - 1; k >= 0; --k) {
            uint32_t m = meta[k];
            const bool was = (m & kMetaWas) != 0, now = ((m >> 16) & 1) != 0;
            if (m & kMetaDead) {
                if (was) contact_event(m, false);
                cm_set(dead, k);
                continue;
            }
            if (!was && now) contact_event(m, true);
            if (was && !now) contact_event(m, false);
            meta[k] = m & 0x0fffffffu;
        }
        compact_contacts(dead);
    }
    MRP_HD void compact_contacts(const CMask& dead) {
        if (cm_any(dead)) {  // rare: compact, preserving order
            int dst = 0;
            for (int k = 0; k < nc; ++k) {
                if (cm_test(dead, k)) continue;
                if (dst != k) {
                    meta[dst] = meta[k];
                    for (int j = 1; j < MRP_CONTACT_WORDS; ++j) g(cw(dst, j)) = g(cw(k, j));
                }
                ++dst;
            }
            nc = dst;
        }
    }
    MRP_HD void collide() {
        CMask dead = cm_none();
        for (int k = nc - 1; k >= 0; --k) {
            uint32_t m = meta[k];
            int fa = m & 0xff, fb = (m >> 8) & 0xff;
            if (!overlap(fat(fa), fat(fb))) {
                if ((m >> 16) & 1) contact_event(m, false);
                cm_set(dead, k);
                continue;
            }
            update_contact(k);
        }
        compact_contacts(dead);
    }

    // ------------------------------------------------------------ broadphase (A.3)
    MRP_HD Box shape_aabb(const float* sh, Xf xf) {  // b2PolygonShape::ComputeAABB
        V2 lo = xmul(xf, sh_v(sh, 0)), hi = lo;
        int cnt = sh_count(sh);
        for (int i = 1; i < cnt; ++i) {
            V2 p = xmul(xf, sh_v(sh, i));
            lo = mk(fmin2(lo.x, p.x), fmin2(lo.y, p.y));
            hi = mk(fmax2(hi.x, p.x), fmax2(hi.y, p.y));
        }
        Box b;
        b.lx = lo.x - kPolygonRadius; b.ly = lo.y - kPolygonRadius;
        b.hx = hi.x + kPolygonRadius; b.hy = hi.y + kPolygonRadius;
        return b;
    }
    // b2Body::SynchronizeFixtures + b2DynamicTree::MoveProxy; returns the moved-proxy mask
    MRP_HD uint32_t synchronize_fixtures(int b, Xf xf1) {
        uint32_t moved = 0;
        Xf xf2 = body_xf(b);
        V2 disp = xf2.p - xf1.p;
        int f0, f1;
        fix_range(b, f0, f1);
        for (int f = f0; f < f1; ++f) {
            const float* sh = fix_shape(f);
            // both b2PolygonShape::ComputeAABB evaluations (at xf1 and xf2) in one rolled vertex loop: same values as two
            // shape_aabb() calls, a fraction of the code (k_post is bound by instruction fetch, profiles/)
            Box a1, a2;
            {
                const V2 v0 = sh_v(sh, 0);
                V2 lo1 = xmul(xf1, v0), hi1 = lo1, lo2 = xmul(xf2, v0), hi2 = lo2;
                const int cnt = sh_count(sh);
#pragma unroll 1
                for (int i = 1; i < cnt; ++i) {
                    const V2 v = sh_v(sh, i);
                    const V2 p1 = xmul(xf1, v), p2 = xmul(xf2, v);
                    lo1 = mk(fmin2(lo1.x, p1.x), fmin2(lo1.y, p1.y));
                    hi1 = mk(fmax2(hi1.x, p1.x), fmax2(hi1.y, p1.y));
                    lo2 = mk(fmin2(lo2.x, p2.x), fmin2(lo2.y, p2.y));
                    hi2 = mk(fmax2(hi2.x, p2.x), fmax2(hi2.y, p2.y));
                }
                a1.lx = lo1.x - kPolygonRadius; a1.ly = lo1.y - kPolygonRadius;
                a1.hx = hi1.x + kPolygonRadius; a1.hy = hi1.y + kPolygonRadius;
                a2.lx = lo2.x - kPolygonRadius; a2.ly = lo2.y - kPolygonRadius;
                a2.hx = hi2.x + kPolygonRadius; a2.hy = hi2.y + kPolygonRadius;
            }
            Box a;
            a.lx = fmin2(a1.lx, a2.lx); a.ly = fmin2(a1.ly, a2.ly);
            a.hx = fmax2(a1.hx, a2.hx); a.hy = fmax2(a1.hy, a2.hy);
            swept[4 * f] = a.lx; swept[4 * f + 1] = a.ly; swept[4 * f + 2] = a.hx; swept[4 * f + 3] = a.hy;
            if (contains(fat(f), a)) continue;
            a.lx = a.lx - kAabbExtension; a.ly = a.ly - kAabbExtension;
            a.hx = a.hx + kAabbExtension; a.hy = a.hy + kAabbExtension;
            V2 d = kAabbMultiplier * disp;
            if (d.x < 0.0f) a.lx += d.x; else a.hx += d.x;
            if (d.y < 0.0f) a.ly += d.y; else a.hy += d.y;
            FA(f, 0) = a.lx; FA(f, 1) = a.ly; FA(f, 2) = a.hx; FA(f, 3) = a.hy;
            moved |= 1u << f;
        }
        return moved;
    }
    MRP_HD bool contact_exists(int fa, int fb) {
        uint32_t key = (uint32_t)fa | ((uint32_t)fb << 8);
        for (int k = 0; k < nc; ++k)
            if ((meta[k] & 0xffffu) == key) return true;
        return false;
    }
    // b2BroadPhase::UpdatePairs + b2ContactManager::AddPair: pairs visited in lexicographic
    // (proxyA < proxyB) order == the sorted pair buffer; each new contact goes to the list head.
    MRP_HD void find_new_contacts(uint32_t moved) {
        if (!moved) return;
        int nf = K.nfix;
        for (int i = 0; i < nf; ++i) {
            bool mi = (moved >> i) & 1;
            uint32_t cand = mi ? (0xffffffffu << (i + 1)) : (moved & (0xffffffffu << (i + 1)));
            if (i + 1 >= 32) cand = 0;
            cand &= (nf >= 32) ? 0xffffffffu : ((1u << nf) - 1u);
            if (!cand) continue;
            int bi = fix_body(i);
            Box ai = fat(i);
            while (cand) {
                int j = ctz32(cand);
                cand &= cand - 1;
                int bj = fix_body(j);
                if (bi == bj) continue;
                if (!is_dyn(bi) && !is_dyn(bj)) continue;
                if (!overlap(ai, fat(j))) continue;
                if (contact_exists(i, j)) continue;
                if (nc >= K.maxc) { overflow = 1; continue; }
                meta[nc] = (uint32_t)i | ((uint32_t)j << 8) | ((uint32_t)bi << 20) | ((uint32_t)bj << 24);
                ++nc;
            }
        }
    }
    static MRP_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
        return __ffs((int)x) - 1;
#else
        return __builtin_ctz(x);
#endif
    }

    // ------------------------------------------------------------ contact solver (A.8)
    MRP_HD float& V(int t, int w) {
        MRP_ASSERT(t >= 0 && t < kMaxC && w >= 0 && w < VC_WORDS, CHK_RECORD);
        return vcp[t * VC_WORDS + w];
    }
    MRP_HD uint32_t vmeta(int t) { return __float_as_uint_(vcp[t * VC_WORDS + VC_META]); }

    // b2ContactSolver ctor + InitializeVelocityConstraints for constraints [0, T)
    MRP_HD void init_constraints(int T, const uint8_t* island_of, bool warm) {
        for (int t = 0; t < T; ++t) {
            int k = order[t];
            uint32_t m = meta[k];
            int bA = (m >> 20) & 15, bB = (m >> 24) & 15;
            int pc = (m >> 18) & 3, type = (m >> 17) & 1;
            int fa = m & 0xff, fb = (m >> 8) & 0xff;
            V(t, VC_FRIC) = sqrtf(ct[CT_FIXFRIC + fa] * ct[CT_FIXFRIC + fb]);
            float mw[12];  // manifold words 2..13 of the contact slot, fetched as one batch of independent loads
#pragma unroll
            for (int i = 0; i < 12; ++i) mw[i] = (i < 4 + 4 * pc) ? gf(cw(k, 2 + i)) : 0.0f;
            V2 ln = mk(mw[0], mw[1]), lp = mk(mw[2], mw[3]);
            V(t, VC_LNX) = ln.x; V(t, VC_LNY) = ln.y; V(t, VC_LPX) = lp.x; V(t, VC_LPY) = lp.y;
            V2 lpt[2];
            float nI[2], tI[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j < pc) {
                    lpt[j] = mk(mw[4 + 4 * j], mw[5 + 4 * j]);
                    nI[j] = warm ? mw[6 + 4 * j] : 0.0f;
                    tI[j] = warm ? mw[7 + 4 * j] : 0.0f;
                } else {
                    lpt[j] = mk(0.0f, 0.0f); nI[j] = 0.0f; tI[j] = 0.0f;
                }
                V(t, VC_LP0X + 2 * j) = lpt[j].x;
                V(t, VC_LP0X + 2 * j + 1) = lpt[j].y;
            }
            float mA = invMass(bA), mB = invMass(bB), iA = invI(bA), iB = invI(bB);
            V2 cA = mk(B(bA, 0), B(bA, 1)), cB = mk(B(bB, 0), B(bB, 1));
            V2 vA = mk(B(bA, 3), B(bA, 4)), vB = mk(B(bB, 3), B(bB, 4));
            float wA = B(bA, 5), wB = B(bB, 5);
            Xf xfA, xfB;
            xfA.q = body_rot(bA, B(bA, 2));
            xfB.q = body_rot(bB, B(bB, 2));
            xfA.p = cA - rmul(xfA.q, localCenter(bA));
            xfB.p = cB - rmul(xfB.q, localCenter(bB));
            // b2WorldManifold::Initialize
            V2 normal, wpt[2];
            if (type == 0) {
                normal = rmul(xfA.q, ln);
                V2 planePoint = xmul(xfA, lp);
                for (int j = 0; j < pc; ++j) {
                    V2 clip = xmul(xfB, lpt[j]);
                    V2 pA = clip + (kPolygonRadius - dot(clip - planePoint, normal)) * normal;
                    V2 pB = clip - kPolygonRadius * normal;
                    wpt[j] = 0.5f * (pA + pB);
                }
            } else {
                normal = rmul(xfB.q, ln);
                V2 planePoint = xmul(xfB, lp);
                for (int j = 0; j < pc; ++j) {
                    V2 clip = xmul(xfA, lpt[j]);
                    V2 pB = clip + (kPolygonRadius - dot(clip - planePoint, normal)) * normal;
                    V2 pA = clip - kPolygonRadius * normal;
                    wpt[j] = 0.5f * (pA + pB);
                }
                normal = -normal;
            }
            V(t, VC_NX) = normal.x; V(t, VC_NY) = normal.y;
            V(t, VC_MA) = mA; V(t, VC_IA) = iA; V(t, VC_MB) = mB; V(t, VC_IB) = iB;
            int vpc = pc;
            V2 tangent = crossVS(normal, 1.0f);
            V2 rAj[2], rBj[2];
            for (int j = 0; j < 2; ++j) {
                float* P = &V(t, VC_PT + 8 * j);
                if (j >= pc) { for (int q = 0; q < 8; ++q) P[q] = 0.0f; continue; }
                V2 rA = wpt[j] - cA, rB = wpt[j] - cB;
                rAj[j] = rA; rBj[j] = rB;
                float rnA = cross(rA, normal), rnB = cross(rB, normal);
                float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                float rtA = cross(rA, tangent), rtB = cross(rB, tangent);
                float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
                P[0] = rA.x; P[1] = rA.y; P[2] = rB.x; P[3] = rB.y;
                P[4] = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
                P[5] = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
                P[6] = nI[j]; P[7] = tI[j];
                // restitution is 0 for every fixture here => velocityBias = -0 * vRel (no effect)
            }
            V(t, VC_K11) = 0.0f; V(t, VC_K12) = 0.0f; V(t, VC_K22) = 0.0f;
            V(t, VC_M11) = 0.0f; V(t, VC_M12) = 0.0f; V(t, VC_M22) = 0.0f;
            if (pc == 2) {
                float rn1A = cross(rAj[0], normal), rn1B = cross(rBj[0], normal);
                float rn2A = cross(rAj[1], normal), rn2B = cross(rBj[1], normal);
                float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
                float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
                float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
                const float k_maxConditionNumber = 1000.0f;
                if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
                    V(t, VC_K11) = k11; V(t, VC_K12) = k12; V(t, VC_K22) = k22;
                    float det = k11 * k22 - k12 * k12;
                    if (det != 0.0f) det = 1.0f / det;
                    V(t, VC_M11) = det * k22; V(t, VC_M12) = -det * k12; V(t, VC_M22) = det * k11;
                } else {
                    vpc = 1;
                }
            }
            (void)vA; (void)vB; (void)wA; (void)wB;
            uint32_t vm = (uint32_t)bA | ((uint32_t)bB << 4) | ((uint32_t)vpc << 8) | ((uint32_t)pc << 10) | ((uint32_t)type << 12) |
                          ((uint32_t)(island_of ? island_of[t] : 0) << 16) | ((uint32_t)k << 24);
            V(t, VC_META) = __uint_as_float_(vm);
        }
    }

    MRP_HD void warm_start(int T) {
        for (int t = 0; t < T; ++t) {
            uint32_t vm = vmeta(t);
            int bA = vm & 15, bB = (vm >> 4) & 15, vpc = (vm >> 8) & 3;
            float mA = invMass(bA), mB = invMass(bB), iA = invI(bA), iB = invI(bB);
            V2 vA = mk(B(bA, 3), B(bA, 4)), vB = mk(B(bB, 3), B(bB, 4));
            float wA = B(bA, 5), wB = B(bB, 5);
            V2 normal = mk(V(t, VC_NX), V(t, VC_NY));
            V2 tangent = crossVS(normal, 1.0f);
            for (int j = 0; j < vpc; ++j) {
                const float* P = &V(t, VC_PT + 8 * j);
                V2 rA = mk(P[0], P[1]), rB = mk(P[2], P[3]);
                V2 Pi = P[6] * normal + P[7] * tangent;
                wA -= iA * cross(rA, Pi);
                vA = vA - mA * Pi;
                wB += iB * cross(rB, Pi);
                vB = vB + mB * Pi;
            }
            B(bA, 3) = vA.x; B(bA, 4) = vA.y; B(bA, 5) = wA;
            B(bB, 3) = vB.x; B(bB, 4) = vB.y; B(bB, 5) = wB;
        }
    }

    // b2ContactSolver::SolveVelocityConstraints x iters, flattened: the (sweep, contact, point) loop nest runs as
    // ONE per-lane loop whose trip is a single point operation — friction point j, then the normal point
    // (1-point manifolds) or the 2-point block solve.  Lanes of a warp advance through their own islands
    // independently, so a warp runs max_lane(total ops) trips instead of 180 x max_lane(contacts) x max(points).
    // The operation order inside an env is exactly Box2D's.  A lane stops after the first sweep that changes
    // nothing: every later sweep would be the identical no-op, so the result equals the full 180 bit for bit.
    MRP_HD void solve_velocity(int T, int iters) {
        if (T == 0) return;
        if (T == 1) {  // register-resident forms (identical arithmetic, see vr_*)
            VelReg r;
            vr_begin(r, 1);
            if (r.vpc == 2) { while (!vr_sweep_single<2>(r, iters)) {} }
            else { while (!vr_sweep_single<1>(r, iters)) {} }
            return;
        }
        if (T == 2) {
            VelReg r0, r1;
            vr_begin_pair(r0, r1);
            while (!vr_sweep_pair(r0, r1, iters)) {}
            return;
        }
        VelReg r;  // larger islands: one contact per trip, the current constraint record in registers
        float imp[4 * kMaxC];
        vr_begin(r, T, imp);
        while (!vr_trip_contact(r, iters, imp)) {}
    }
    // ---- register-resident form of the velocity solve (solver kernel): the current contact's constraint record and
    // the velocities of its two bodies live in registers while its point operations run; they are exchanged with
    // memory only when the lane moves on to another contact.  Islands with one contact (the common case) therefore
    // iterate entirely in registers.  Arithmetic and operation order are Box2D's (friction points, then the normal point / block solve, contact by contact).
    struct VelReg {
        int T, t, j, sweep, bA, bB, vpc;
        bool changed;
        float mA, iA, mB, iB, fric;
        V2 n, rA0, rB0, rA1, rB1;
        float nM0, tM0, nI0, tI0, nM1, tM1, nI1, tI1;
        float k11, k12, k22, m11, m12, m22;
        V2 vA, vB;
        float wA, wB;
    };
    // imp != NULL: the accumulated impulses of the island live in a lane-local array during the sweeps (vr_imp_*), so the
    // constraint records are read-only while the island iterates and stay resident in L1
    MRP_HD void vr_load(VelReg& r, const float* imp = nullptr) {
        const float* C = &V(r.t, 0);
        const uint32_t vm = __float_as_uint_(C[VC_META]);
        r.bA = vm & 15; r.bB = (vm >> 4) & 15; r.vpc = (vm >> 8) & 3;
        r.mA = C[VC_MA]; r.iA = C[VC_IA]; r.mB = C[VC_MB]; r.iB = C[VC_IB];
        r.fric = C[VC_FRIC];
        r.n = mk(C[VC_NX], C[VC_NY]);
        r.rA0 = mk(C[VC_PT + 0], C[VC_PT + 1]); r.rB0 = mk(C[VC_PT + 2], C[VC_PT + 3]);
        r.nM0 = C[VC_PT + 4]; r.tM0 = C[VC_PT + 5];
        r.rA1 = mk(C[VC_PT + 8], C[VC_PT + 9]); r.rB1 = mk(C[VC_PT + 10], C[VC_PT + 11]);
        r.nM1 = C[VC_PT + 12]; r.tM1 = C[VC_PT + 13];
        if (imp) { const float* I = imp + 4 * r.t; r.nI0 = I[0]; r.tI0 = I[1]; r.nI1 = I[2]; r.tI1 = I[3]; }
        else { r.nI0 = C[VC_PT + 6]; r.tI0 = C[VC_PT + 7]; r.nI1 = C[VC_PT + 14]; r.tI1 = C[VC_PT + 15]; }
        r.k11 = C[VC_K11]; r.k12 = C[VC_K12]; r.k22 = C[VC_K22];
        r.m11 = C[VC_M11]; r.m12 = C[VC_M12]; r.m22 = C[VC_M22];
        vr_load_vel(r);
    }
    MRP_HD void vr_load_vel(VelReg& r) {
        const float* pA = bp(r.bA);
        const float* pB = bp(r.bB);
        r.vA = mk(pA[3 * MRP_SS], pA[4 * MRP_SS]); r.wA = pA[5 * MRP_SS];
        r.vB = mk(pB[3 * MRP_SS], pB[4 * MRP_SS]); r.wB = pB[5 * MRP_SS];
    }
    MRP_HD void vr_store_vel(const VelReg& r) {
        float* pA = bp(r.bA);
        float* pB = bp(r.bB);
        pA[3 * MRP_SS] = r.vA.x; pA[4 * MRP_SS] = r.vA.y; pA[5 * MRP_SS] = r.wA;
        pB[3 * MRP_SS] = r.vB.x; pB[4 * MRP_SS] = r.vB.y; pB[5 * MRP_SS] = r.wB;
    }
    MRP_HD void vr_store(const VelReg& r, float* imp = nullptr) {
        if (imp) { float* I = imp + 4 * r.t; I[0] = r.nI0; I[1] = r.tI0; I[2] = r.nI1; I[3] = r.tI1; }
        else { float* C = &V(r.t, 0); C[VC_PT + 6] = r.nI0; C[VC_PT + 7] = r.tI0; C[VC_PT + 14] = r.nI1; C[VC_PT + 15] = r.tI1; }
        float* pA = bp(r.bA);
        float* pB = bp(r.bB);
        pA[3 * MRP_SS] = r.vA.x; pA[4 * MRP_SS] = r.vA.y; pA[5 * MRP_SS] = r.wA;
        pB[3 * MRP_SS] = r.vB.x; pB[4 * MRP_SS] = r.vB.y; pB[5 * MRP_SS] = r.wB;
    }
    MRP_HD void vr_begin(VelReg& r, int T, float* imp = nullptr) {
        r.T = T; r.t = 0; r.j = 0; r.sweep = 0; r.changed = false;
        if (imp)   // records -> lane-local impulses
            for (int t = 0; t < T; ++t) {
                const float* C = &V(t, 0);
                imp[4 * t] = C[VC_PT + 6]; imp[4 * t + 1] = C[VC_PT + 7]; imp[4 * t + 2] = C[VC_PT + 14]; imp[4 * t + 3] = C[VC_PT + 15];
            }
        vr_load(r, imp);
    }
    MRP_HD void vr_imp_flush(const float* imp, int T) {  // lane-local impulses -> records (StoreImpulses reads them there)
        for (int t = 0; t < T; ++t) {
            float* C = &V(t, 0);
            C[VC_PT + 6] = imp[4 * t]; C[VC_PT + 7] = imp[4 * t + 1]; C[VC_PT + 14] = imp[4 * t + 2]; C[VC_PT + 15] = imp[4 * t + 3];
        }
    }
    // one point operation; returns true when the solve is finished (everything written back)
    // the two kinds of point operation on the register-resident contact
    template <int KIND>  // 0: friction point 0, 1: friction point 1, 2: normal point 0 (1-point manifolds)
    MRP_HD void vr_op_1d(VelReg& r) {
        constexpr bool fr = KIND != 2, second = KIND == 1;
        const V2 dir = fr ? crossVS(r.n, 1.0f) : r.n;
        const V2 rA = second ? r.rA1 : r.rA0, rB = second ? r.rB1 : r.rB0;
        const V2 dv = r.vB + crossSV(r.wB, rB) - r.vA - crossSV(r.wA, rA);
        const float vd = dot(dv, dir);
        const float acc = fr ? (second ? r.tI1 : r.tI0) : r.nI0;
        const float mass = fr ? (second ? r.tM1 : r.tM0) : r.nM0;
        float lambda = mass * (-vd);  // tangentMass*(-vt)  ==  -normalMass*vn (exact sign symmetry)
        const float maxFriction = r.fric * (second ? r.nI1 : r.nI0);
        const float newImpulse = fr ? clampf(acc + lambda, -maxFriction, maxFriction) : fmax2(acc + lambda, 0.0f);
        lambda = newImpulse - acc;
        if (!fr) r.nI0 = newImpulse;
        else if (second) r.tI1 = newImpulse;
        else r.tI0 = newImpulse;
        r.changed = r.changed || (lambda != 0.0f);
        const V2 Pi = lambda * dir;
        r.vA = r.vA - r.mA * Pi;
        r.wA -= r.iA * cross(rA, Pi);
        r.vB = r.vB + r.mB * Pi;
        r.wB += r.iB * cross(rB, Pi);
    }
    MRP_HD void vr_op_block(VelReg& r) {
        const float ax = r.nI0, ay = r.nI1;
        const V2 dv1 = r.vB + crossSV(r.wB, r.rB0) - r.vA - crossSV(r.wA, r.rA0);
        const V2 dv2 = r.vB + crossSV(r.wB, r.rB1) - r.vA - crossSV(r.wA, r.rA1);
        float vn1 = dot(dv1, r.n), vn2 = dot(dv2, r.n);
        float bx = vn1, by = vn2;
        bx -= r.k11 * ax + r.k12 * ay;
        by -= r.k12 * ax + r.k22 * ay;
        float xx, xy;
        bool ok = false;
        xx = -(r.m11 * bx + r.m12 * by);
        xy = -(r.m12 * bx + r.m22 * by);
        if (xx >= 0.0f && xy >= 0.0f) ok = true;
        if (!ok) {
            xx = -r.nM0 * bx; xy = 0.0f;
            vn2 = r.k12 * xx + by;
            if (xx >= 0.0f && vn2 >= 0.0f) ok = true;
        }
        if (!ok) {
            xx = 0.0f; xy = -r.nM1 * by;
            vn1 = r.k12 * xy + bx;
            if (xy >= 0.0f && vn1 >= 0.0f) ok = true;
        }
        if (!ok) {
            xx = 0.0f; xy = 0.0f;
            if (bx >= 0.0f && by >= 0.0f) ok = true;
        }
        if (ok) {
            const float dx = xx - ax, dy = xy - ay;
            const V2 Pa = dx * r.n, Pb = dy * r.n;
            r.vA = r.vA - r.mA * (Pa + Pb);
            r.wA -= r.iA * (cross(r.rA0, Pa) + cross(r.rA1, Pb));
            r.vB = r.vB + r.mB * (Pa + Pb);
            r.wB += r.iB * (cross(r.rB0, Pa) + cross(r.rB1, Pb));
            r.nI0 = xx; r.nI1 = xy;
            r.changed = r.changed || (dx != 0.0f) || (dy != 0.0f);
        }
    }
    // generic form for the persistent solver: one whole CONTACT per trip (friction points, then the normal point or the
    // block solve — Box2D's order), so lanes of a warp only diverge between 1- and 2-point manifolds instead of
    // between four kinds of point operation, and the per-trip bookkeeping is paid once per contact
    MRP_HD bool vr_trip_contact(VelReg& r, int iters, float* imp = nullptr) {
        vr_contact_ops(r);
        bool wrapped = true;
        if (r.T > 1) {
            vr_store(r, imp);
            wrapped = ++r.t == r.T;
            if (wrapped) r.t = 0;
            vr_load(r, imp);
        }
        if (wrapped) {
            ++r.sweep;
            if (!r.changed || r.sweep == iters) {
                if (r.T == 1) vr_store(r);
                else if (imp) vr_imp_flush(imp, r.T);
                return true;
            }
            r.changed = false;
        }
        return false;
    }
    // two-contact islands: both constraint records and the velocities of their bodies stay in registers; the bodies the two
    // contacts share (one or both) are handed from one contact to the other by vr_forward
    MRP_HD void vr_begin_pair(VelReg& r0, VelReg& r1) {
        r0.T = 2; r0.t = 0; r0.j = 0; r0.sweep = 0; r0.changed = false;
        vr_load(r0);
        r1.T = 2; r1.t = 1; r1.j = 0; r1.sweep = 0; r1.changed = false;
        vr_load(r1);
    }
    MRP_HD void vr_contact_ops(VelReg& r) {
        vr_op_1d<0>(r);
        if (r.vpc == 2) { vr_op_1d<1>(r); vr_op_block(r); }
        else vr_op_1d<2>(r);
    }
    // the velocities of the bodies contact d shares with contact s, straight from s's registers: what a store of s's velocities to the
    // lane's body slots followed by a load of d's would deliver (a body d does not share with s still holds what d left there)
    MRP_HD void vr_forward(const VelReg& s, VelReg& d) {
        if (d.bA == s.bA) { d.vA = s.vA; d.wA = s.wA; } else if (d.bA == s.bB) { d.vA = s.vB; d.wA = s.wB; }
        if (d.bB == s.bA) { d.vB = s.vA; d.wB = s.wA; } else if (d.bB == s.bB) { d.vB = s.vB; d.wB = s.wB; }
    }
    MRP_HD bool vr_sweep_pair(VelReg& r0, VelReg& r1, int iters) {
        // r0 holds current velocities on entry (vr_begin_pair, or the forward at the end of the previous sweep): the shared-memory
        // round trip between the two contacts (12 stores + 12 loads on the dependent chain of every sweep) is replaced by selects
        vr_contact_ops(r0);
        vr_forward(r0, r1);
        vr_contact_ops(r1);
        ++r0.sweep;
        if (!(r0.changed || r1.changed) || r0.sweep == iters) {
            vr_store(r0);
            vr_store(r1);   // contact 1 ran last: its velocities are the final ones for shared bodies
            return true;
        }
        vr_forward(r1, r0);
        r0.changed = false;
        r1.changed = false;
        return false;
    }

    // single-contact islands: one trip = one whole sweep, branch-free op sequence, nothing leaves the registers.
    // VPC = 1: friction, normal.  VPC = 2: friction, friction, block solve.
    template <int VPC>
    MRP_HD bool vr_sweep_single(VelReg& r, int iters) {
        vr_op_1d<0>(r);
        if (VPC == 2) { vr_op_1d<1>(r); vr_op_block(r); }
        else vr_op_1d<2>(r);
        ++r.sweep;
        if (!r.changed || r.sweep == iters) {
            vr_store(r);
            return true;
        }
        r.changed = false;
        return false;
    }

    MRP_HD void store_impulses(int T) {
        for (int t = 0; t < T; ++t) {
            uint32_t vm = vmeta(t);
            int vpc = (vm >> 8) & 3, k = (vm >> 24) & 0xff;
            for (int j = 0; j < vpc; ++j) {
                gsf(cw(k, 8 + 4 * j), V(t, VC_PT + 8 * j + 6));
                gsf(cw(k, 9 + 4 * j), V(t, VC_PT + 8 * j + 7));
            }
        }
    }
    MRP_HD void integrate_position(int b, float h) {
        V2 v = mk(B(b, 3), B(b, 4));
        float w = B(b, 5);
        V2 tr = h * v;
        if (dot(tr, tr) > kMaxTranslationSquared) {
            float ratio = kMaxTranslation / length(tr);
            v = ratio * v;
        }
        float rotation = h * w;
        if (rotation * rotation > kMaxRotationSquared) {
            float ratio = kMaxRotation / fabsf(rotation);
            w *= ratio;
        }
        B(b, 0) += h * v.x;
        B(b, 1) += h * v.y;
        B(b, 2) += h * w;
        B(b, 3) = v.x; B(b, 4) = v.y; B(b, 5) = w;
    }
    // b2ContactSolver::SolvePositionConstraints (toiA < 0) / SolveTOIPositionConstraints, up to maxSweeps sweeps,
    // flattened like solve_velocity: one trip = one manifold point.  Islands are solved "in parallel" inside one env:
    // an island whose sweep ended with minSeparation >= limit is finished (Box2D breaks out of its loop) and its
    // constraints are skipped from then on.  Per-constraint min separations are folded into per-island flags.
    struct PosState {
        uint32_t done, bad;
        int t, j, sweep;
        float minSep;
    };
    MRP_HD void pos_begin(PosState& st) { st.done = 0; st.bad = 0; st.t = 0; st.j = 0; st.sweep = 0; st.minSep = 0.0f; }
    MRP_HD void solve_position(int T, int maxSweeps, int toiA, int toiB) {
        if (T == 0) return;
        PosState st;
        pos_begin(st);
        while (!pos_trip(st, T, maxSweeps, toiA, toiB)) {}
    }
    // one CONTACT per trip (all its manifold points, in order): the constraint record and the two bodies are read once,
    // corrected in registers and written back once
    template <bool INL = false>
    MRP_HD bool pos_trip(PosState& st, int T, int maxSweeps, int toiA, int toiB) {
        const bool toi = toiA >= 0;
        const float baum = toi ? kToiBaumgarte : kBaumgarte;
        const float lim = toi ? -1.5f * kLinearSlop : -3.0f * kLinearSlop;
        uint32_t done = st.done, bad = st.bad;
        int t = st.t;
        float minSep = st.minSep;
        bool finished = false;
        {
            const uint32_t vm = vmeta(t);
            const int isl = (vm >> 16) & 0xff;
            const int ppc = (vm >> 10) & 3;
            if (!((done >> isl) & 1)) {
                stat_pos_pts += (uint32_t)ppc;
                const int bA = vm & 15, bB = (vm >> 4) & 15, type = (vm >> 12) & 1;
                float mA = V(t, VC_MA), mB = V(t, VC_MB), iA = V(t, VC_IA), iB = V(t, VC_IB);
                if (toi) {
                    if (bA != toiA && bA != toiB) { mA = 0.0f; iA = 0.0f; }
                    if (bB != toiA && bB != toiB) { mB = 0.0f; iB = 0.0f; }
                }
                V2 cA = mk(B(bA, 0), B(bA, 1)), cB = mk(B(bB, 0), B(bB, 1));
                float aA = B(bA, 2), aB = B(bB, 2);
                const V2 ln = mk(V(t, VC_LNX), V(t, VC_LNY)), lp = mk(V(t, VC_LPX), V(t, VC_LPY));
                const V2 lp0 = mk(V(t, VC_LP0X), V(t, VC_LP0Y)), lp1 = mk(V(t, VC_LP1X), V(t, VC_LP1Y));
#pragma unroll 1
                for (int j = 0; j < ppc; ++j) {
                    Xf xfA, xfB;
                    xfA.q = body_rot<INL>(bA, aA);
                    xfB.q = body_rot<INL>(bB, aB);
                    xfA.p = cA - rmul(xfA.q, localCenter(bA));
                    xfB.p = cB - rmul(xfB.q, localCenter(bB));
                    const V2 lpj = j == 0 ? lp0 : lp1;
                    V2 normal, point;
                    float separation;
                    if (type == 0) {
                        normal = rmul(xfA.q, ln);
                        V2 planePoint = xmul(xfA, lp);
                        V2 clip = xmul(xfB, lpj);
                        separation = dot(clip - planePoint, normal) - kPolygonRadius - kPolygonRadius;
                        point = clip;
                    } else {
                        normal = rmul(xfB.q, ln);
                        V2 planePoint = xmul(xfB, lp);
                        V2 clip = xmul(xfA, lpj);
                        separation = dot(clip - planePoint, normal) - kPolygonRadius - kPolygonRadius;
                        point = clip;
                        normal = -normal;
                    }
                    const V2 rA = point - cA, rB = point - cB;
                    minSep = fmin2(minSep, separation);
                    const float C = clampf(baum * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
                    const float rnA = cross(rA, normal), rnB = cross(rB, normal);
                    const float Kn = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                    const float impulse = Kn > 0.0f ? -C / Kn : 0.0f;
                    const V2 Pi = impulse * normal;
                    cA = cA - mA * Pi;
                    aA -= iA * cross(rA, Pi);
                    cB = cB + mB * Pi;
                    aB += iB * cross(rB, Pi);
                }
                B(bA, 0) = cA.x; B(bA, 1) = cA.y; B(bA, 2) = aA;
                B(bB, 0) = cB.x; B(bB, 1) = cB.y; B(bB, 2) = aB;
            }
            if (!(minSep >= lim)) bad |= 1u << isl;
            minSep = 0.0f;
            if (++t == T) {
                t = 0;
                ++st.sweep;
                done = ~bad;
                if (!bad || st.sweep == maxSweeps) finished = true;
                bad = 0;
            }
        }
        st.done = done; st.bad = bad; st.t = t; st.j = 0; st.minSep = minSep;
        return finished;
    }

    // ------------------------------------------------------------ b2World::Solve (A.7, A.8)
    // island build (A.7): fills order[] (solver order = DFS discovery order) and island_of[]; returns T
    MRP_HD int build_islands(uint8_t* island_of) {
        // island build: DFS seeds in body-list order (newest first): agent n-1 .. agent 0, block
        int T = 0;
        CMask touch = cm_none();
        for (int k = 0; k < nc; ++k) if ((meta[k] >> 16) & 1u) cm_set(touch, k);
        if (cm_single(touch)) {  // a single touching contact is its own island: no search needed
            order[0] = (uint8_t)cm_first(touch);
            island_of[0] = 0;
            return 1;
        }
#if MRP_MAXC == 32
        if (touch) {
            // per-body sets of touching contacts, so that popping a body visits only its own contacts (newest first)
            // instead of scanning the whole contact list; same visiting order, hence the same solver order
            uint32_t bmask[16];
            for (int b = 0; b < K.nb + 4; ++b) bmask[b] = 0u;
            for (uint32_t rest = touch; rest; rest &= rest - 1u) {
                const int k = ctz_u32(rest);
                const uint32_t m = meta[k];
                bmask[(m >> 20) & 15] |= 1u << k;
                bmask[(m >> 24) & 15] |= 1u << k;
            }
            uint32_t bflag = 0, cflag = 0;
            int nisl = 0;
            for (int seed = K.nb - 1; seed >= 0; --seed) {
                if ((bflag >> seed) & 1) continue;
                if (!bmask[seed]) { bflag |= 1u << seed; continue; }  // no touching contact: an island without constraints
                uint64_t stack = (uint64_t)seed;  // 4 bits per entry
                int sp = 1;
                bflag |= 1u << seed;
                uint32_t statics = 0;
                const int T0 = T;
                while (sp > 0) {
                    --sp;
                    const int b = (int)((stack >> (4 * sp)) & 15u);
                    if (!is_dyn(b)) { statics |= 1u << b; continue; }
                    for (uint32_t cand = bmask[b] & ~cflag; cand;) {
                        const int k = 31 - clz_u32(cand);
                        cand &= ~(1u << k);
                        const uint32_t m = meta[k];
                        const int bA = (m >> 20) & 15, bB = (m >> 24) & 15;
                        order[T] = (uint8_t)k;
                        island_of[T] = (uint8_t)nisl;
                        ++T;
                        cflag |= 1u << k;
                        const int other = bA == b ? bB : bA;
                        if ((bflag >> other) & 1) continue;
                        stack = (stack & ~(15ull << (4 * sp))) | ((uint64_t)other << (4 * sp));
                        ++sp;
                        bflag |= 1u << other;
                    }
                }
                bflag &= ~statics;  // static bodies may join later islands
                if (T > T0) ++nisl;
            }
        }
#else
        if (cm_any(touch)) {
            uint32_t bflag = 0;
            CMask cflag = cm_none();
            int nisl = 0;
            for (int seed = K.nb - 1; seed >= 0; --seed) {
                if ((bflag >> seed) & 1) continue;
                uint64_t stack = (uint64_t)seed;  // 4 bits per entry
                int sp = 1;
                bflag |= 1u << seed;
                uint32_t statics = 0;
                int T0 = T;
                while (sp > 0) {
                    --sp;
                    int b = (int)((stack >> (4 * sp)) & 15u);
                    if (!is_dyn(b)) { statics |= 1u << b; continue; }
                    for (int k = nc - 1; k >= 0; --k) {
                        if (!cm_test(touch, k) || cm_test(cflag, k)) continue;
                        uint32_t m = meta[k];
                        int bA = (m >> 20) & 15, bB = (m >> 24) & 15;
                        if (bA != b && bB != b) continue;
                        order[T] = (uint8_t)k;
                        island_of[T] = (uint8_t)nisl;
                        ++T;
                        cm_set(cflag, k);
                        int other = bA == b ? bB : bA;
                        if ((bflag >> other) & 1) continue;
                        stack = (stack & ~(15ull << (4 * sp))) | ((uint64_t)other << (4 * sp));
                        ++sp;
                        bflag |= 1u << other;
                    }
                }
                bflag &= ~statics;  // static bodies may join later islands
                if (T > T0) ++nisl;
            }
        }
#endif
        return T;
    }

    MRP_HD void solve_islands() {
        uint8_t island_of[kMaxC];
        int T = build_islands(island_of);
        init_constraints(T, island_of, true);
        warm_start(T);
        solve_velocity(T, 180);
        store_impulses(T);
        for (int b = 0; b < K.nb; ++b) integrate_position(b, K.h);
        solve_position(T, 60, -1, -1);
    }

    // ------------------------------------------------------------ b2World::SolveTOI (A.10)
    MRP_HD Sweep body_sweep(int b) {
        Sweep s;
        if (b < K.nb) {
            s.lc = localCenter(b);
            s.c0 = mk(BX(b, c0f), BX(b, c0f + 1));
            s.a0 = BX(b, c0f + 2);
            s.c = mk(B(b, 0), B(b, 1));
            s.a = B(b, 2);
        } else {
            s.lc = mk(0.0f, 0.0f);
            s.c0 = wall_pos(b);
            s.c = s.c0;
            s.a0 = 0.0f;
            s.a = 0.0f;
        }
        return s;
    }
    MRP_HD void sweep_advance(int b, float alpha) {  // b2Sweep::Advance (on c0/a0/alpha0 only)
        float& al0 = alpha0(b);
        if (b < K.nb) {
            float beta = (alpha - al0) / (1.0f - al0);
            BX(b, c0f) += beta * (B(b, 0) - BX(b, c0f));
            BX(b, c0f + 1) += beta * (B(b, 1) - BX(b, c0f + 1));
            BX(b, c0f + 2) += beta * (B(b, 2) - BX(b, c0f + 2));
        }
        al0 = alpha;
    }
    MRP_HD void body_advance(int b, float alpha) {  // b2Body::Advance
        sweep_advance(b, alpha);
        if (b < K.nb) {
            B(b, 0) = BX(b, c0f); B(b, 1) = BX(b, c0f + 1); B(b, 2) = BX(b, c0f + 2);
            sync_transform(b);
        }
    }

    MRP_HDN void toi_event(int minK, float minAlpha, CMask& toiFlag, CMask& enabled) {
        uint32_t m = meta[minK];
        int bA = (m >> 20) & 15, bB = (m >> 24) & 15;
        // backups of the two sweeps
        float bk[2][7];
        int two[2] = {bA, bB};
        for (int s = 0; s < 2; ++s) {
            int b = two[s];
            if (b < K.nb) {
                bk[s][0] = BX(b, c0f); bk[s][1] = BX(b, c0f + 1); bk[s][2] = BX(b, c0f + 2);
                bk[s][3] = B(b, 0); bk[s][4] = B(b, 1); bk[s][5] = B(b, 2);
            }
            bk[s][6] = alpha0(b);
        }
        body_advance(bA, minAlpha);
        body_advance(bB, minAlpha);
        update_contact(minK);
        cm_set(enabled, minK);
        cm_clr(toiFlag, minK);
        ++toiCount[minK];
        if (!((meta[minK] >> 16) & 1)) {
            cm_clr(enabled, minK);
            for (int s = 0; s < 2; ++s) {
                int b = two[s];
                if (b < K.nb) {
                    BX(b, c0f) = bk[s][0]; BX(b, c0f + 1) = bk[s][1]; BX(b, c0f + 2) = bk[s][2];
                    B(b, 0) = bk[s][3]; B(b, 1) = bk[s][4]; B(b, 2) = bk[s][5];
                    sync_transform(b);
                }
                alpha0(b) = bk[s][6];
            }
            return;
        }
        // mini island: bA, bB, minContact + the dynamic body's other touching static contacts
        int T = 0;
        order[T++] = (uint8_t)minK;
        CMask cflag = cm_none();
        cm_set(cflag, minK);
        uint32_t bflag = (1u << bA) | (1u << bB);
        for (int s = 0; s < 2; ++s) {
            int body = two[s];
            if (!is_dyn(body)) continue;
            for (int k = nc - 1; k >= 0; --k) {
                if (T == kMaxC) break;
                if (cm_test(cflag, k)) continue;
                uint32_t mk_ = meta[k];
                int cA = (mk_ >> 20) & 15, cB = (mk_ >> 24) & 15;
                if (cA != body && cB != body) continue;
                int other = cA == body ? cB : cA;
                if (is_dyn(other)) continue;  // no bullets
                float backupAlpha = alpha0(other);
                if (!((bflag >> other) & 1)) body_advance(other, minAlpha);
                update_contact(k);
                cm_set(enabled, k);
                if (!((meta[k] >> 16) & 1)) { alpha0(other) = backupAlpha; continue; }
                cm_set(cflag, k);
                order[T++] = (uint8_t)k;
                bflag |= 1u << other;
            }
        }
        float subdt = (1.0f - minAlpha) * K.h;
        // b2Island::SolveTOI
        init_constraints(T, nullptr, false);
        solve_position(T, 20, bA, bB);
        for (int s = 0; s < 2; ++s) {
            int b = two[s];
            if (b < K.nb) { BX(b, c0f) = B(b, 0); BX(b, c0f + 1) = B(b, 1); BX(b, c0f + 2) = B(b, 2); }
        }
        init_constraints(T, nullptr, false);
        solve_velocity(T, 180);
        uint32_t moved = 0;
        for (int b = 0; b < K.nb; ++b) {
            if (!((bflag >> b) & 1)) continue;
            integrate_position(b, subdt);
            sync_transform(b);
        }
        for (int b = 0; b < K.nb; ++b) {
            if (!((bflag >> b) & 1)) continue;
            Xf xf1;
            xf1.q = rot_set(BX(b, c0f + 2));
            xf1.p = mk(BX(b, c0f), BX(b, c0f + 1)) - rmul(xf1.q, localCenter(b));
            moved |= synchronize_fixtures(b, xf1);
            for (int k = 0; k < nc; ++k) {
                uint32_t mk_ = meta[k];
                if ((int)((mk_ >> 20) & 15) == b || (int)((mk_ >> 24) & 15) == b) cm_clr(toiFlag, k);
            }
        }
        int nc0 = nc;
        find_new_contacts(moved);
        for (int k = nc0; k < nc; ++k) { toi[k] = 1.0f; toiCount[k] = 0; cm_set(enabled, k); cm_clr(toiFlag, k); }
    }

    // Exact TOI culling.  b2TimeOfImpact can only report e_touching (the one state that yields alpha < 1) if the
    // distance between the two core polygons drops below target + tolerance = 0.00625 somewhere on the sweep: every
    // separation it evaluates along its axis is bounded below by the true distance at that time.
    // The swept AABB of the fixture (vertices at both ends of the sweep) minus a curvature term for the rotation in
    // between bounds that distance from below along the wall's axis-aligned faces; when even this bound stays above
    // 0.00625 (+ guard) the call is skipped and alpha = 1, exactly what the full algorithm would return.
    MRP_HD bool toi_provably_one(int f, int b, int wall) {
        const float da = B(b, 2) - BX(b, c0f + 2);
        if (!(fabsf(da) < 0.5f)) return false;
        const float* sx = shape_x(f);
        const V2 lc = localCenter(b);
        const float arm = sx[4] + length(mk(sx[0] - lc.x, sx[1] - lc.y));
        const float curv = 0.25f * arm * da * da;
        const float* w = ct + CT_WALLBOX + 4 * wall;
        const float* a = swept + 4 * f;
        const float r = kPolygonRadius;
        const float gx = fmax2(w[0] - (a[2] - r), (a[0] + r) - w[2]);
        const float gy = fmax2(w[1] - (a[3] - r), (a[1] + r) - w[3]);
        return fmax2(gx, gy) - curv > 0.00625f + 0.001f;
    }

    // returns false when a TOI event must be processed but allow_events is false (the caller then defers this env
    // to the event kernel, which redoes the scan from the same state)
    // First TOI scan of a step in k_post, where events are deferred to k_post_events: every alpha0 is still 0 and nothing
    // has been advanced, so "no event" is certain when every wall contact is culled by the swept-AABB bound.  An env with
    // a contact the bound cannot clear is handed to the event kernel, which runs the full solve_toi() (b2TimeOfImpact /
    // GJK included) from the same state: k_post itself never evaluates a TOI — those calls ran with ~1 active lane per
    // warp and their code (3 k instructions) no longer sits in the hot kernel.
    MRP_HD bool toi_scan_needs_events() {
#pragma unroll 1
        for (int k = nc - 1; k >= 0; --k) {
            const uint32_t m = meta[k];
            const int bA = (m >> 20) & 15, bB = (m >> 24) & 15;
            if (is_dyn(bA) && is_dyn(bB)) continue;
            // fixture A is the dynamic one (walls are the last fixtures)
            if (!toi_provably_one((int)(m & 0xff), bA, bB - K.nb)) return true;
        }
        return false;
    }

    MRP_HD bool solve_toi(bool allow_events = true) {
        if (!allow_events) return !toi_scan_needs_events();
        bool wallc = false;  // contacts with a static body: the only TOI candidates (no bullets)
#pragma unroll 1
        for (int k = 0; k < nc; ++k) {
            uint32_t m = meta[k];
            if (!is_dyn((m >> 20) & 15) || !is_dyn((m >> 24) & 15)) wallc = true;
        }
        if (!wallc) return true;
        alpha_none = 0.0f;
        for (int b = 0; fdyn != 11 && b < K.nb; ++b) BX(b, c0f + 3) = 0.0f;
        for (int k = 0; k < 4; ++k) wallAlpha0[k] = 0.0f;
#pragma unroll 1
        for (int k = 0; k < nc; ++k) { toi[k] = 1.0f; toiCount[k] = 0; }
        CMask toiFlag = cm_none(), enabled = cm_all();
        for (;;) {
            int minK = -1;
            float minAlpha = 1.0f;
            // The scan walks the contacts newest first, as b2World::SolveTOI does.  It is written as "walk down to the next
            // contact that needs a b2TimeOfImpact evaluation (cheap), then evaluate it" so that the lanes of a warp make their
            // n-th evaluation in the same trip of the outer loop: with the evaluation inside the contact loop every lane called
            // it at its own trip count and the warp ran the (3 k instruction) calls one lane after the other.
            int k = nc - 1;
            for (;;) {
                int kt = -1;
                float al0 = 0.0f;
#pragma unroll 1
                for (; k >= 0; --k) {
                    if (!cm_test(enabled, k)) continue;
                    if (toiCount[k] > kMaxSubSteps) continue;
                    float alpha;
                    if (cm_test(toiFlag, k)) {
                        alpha = toi[k];
                    } else {
                        const uint32_t m = meta[k];
                        const int bA = (m >> 20) & 15, bB = (m >> 24) & 15;
                        if (is_dyn(bA) && is_dyn(bB)) continue;
                        // b2World::SolveTOI first aligns the two sweeps to the larger alpha0 (a side effect that later contacts
                        // of the same bodies see, walls included), then calls b2TimeOfImpact
                        al0 = alpha0(bA);
                        if (alpha0(bA) < alpha0(bB)) {
                            al0 = alpha0(bB);
                            sweep_advance(bA, al0);
                        } else if (alpha0(bB) < alpha0(bA)) {
                            al0 = alpha0(bA);
                            sweep_advance(bB, al0);
                        }
                        // fixture A is the dynamic one (walls are the last fixtures).  The culling bound covers every
                        // intermediate pose of the sweep swept[] was computed for, hence also an advanced remainder of it
                        if (!toi_provably_one((int)(m & 0xff), bA, bB - K.nb)) { kt = k; break; }
                        alpha = 1.0f;
                        toi[k] = 1.0f;
                        cm_set(toiFlag, k);
                    }
                    if (alpha < minAlpha) { minK = k; minAlpha = alpha; }
                }
                if (kt < 0) break;
                {
                    const uint32_t m = meta[kt];
                    const int bA = (m >> 20) & 15, bB = (m >> 24) & 15;
                    float t;
                    ++stat_toi;
#ifdef MRP_TAILPROBE
                    const long long tp_c0 = tp_clock();
#endif
                    const int st = time_of_impact(&t, fix_shape(m & 0xff), body_sweep(bA), fix_shape((m >> 8) & 0xff), body_sweep(bB));
#ifdef MRP_TAILPROBE
                    tp_toi_clk += tp_clock() - tp_c0;
#endif
                    const float alpha = st == kToiTouching ? fmin2(al0 + (1.0f - al0) * t, 1.0f) : 1.0f;
                    toi[kt] = alpha;
                    cm_set(toiFlag, kt);
                    if (alpha < minAlpha) { minK = kt; minAlpha = alpha; }
                }
                k = kt - 1;
            }
            if (minK < 0 || 1.0f - 10.0f * kEps < minAlpha) break;
            if (!allow_events) return false;
#ifdef MRP_TAILPROBE
            const long long tp_c1 = tp_clock();
#endif
            toi_event(minK, minAlpha, toiFlag, enabled);
#ifdef MRP_TAILPROBE
            tp_evt_clk += tp_clock() - tp_c1;
            ++tp_evt_n;
#endif
        }
        return true;
    }

    // ------------------------------------------------------------ b2World::Step (A.6)
    MRP_HD void world_step(bool new_fixtures) {
        if (new_fixtures) find_new_contacts(0xffffffffu);
        collide();
        // xf1 of SynchronizeFixtures == the transform the step started with (c0, a0 are set from c, a)
        for (int b = 0; b < K.nb; ++b) { BX(b, c0f) = B(b, 0); BX(b, c0f + 1) = B(b, 1); BX(b, c0f + 2) = B(b, 2); }
        solve_islands();
        post_solve(true);
    }
    // tail of b2World::Solve + SolveTOI: SynchronizeFixtures for every dynamic body, FindNewContacts, TOI
    MRP_HD bool post_solve(bool allow_events) {
        uint32_t moved = 0;
        for (int b = K.nb - 1; b >= 0; --b) {
            // xf1 = transform at (c0, a0); its rotation is the one the step started with
            Xf xf1;
            xf1.q.s = BX(b, 6);
            xf1.q.c = BX(b, 7);
            xf1.p = mk(BX(b, c0f), BX(b, c0f + 1)) - rmul(xf1.q, localCenter(b));
            sync_transform(b);
            moved |= synchronize_fixtures(b, xf1);
        }
        find_new_contacts(moved);
        return solve_toi(allow_events);
    }
};

// k_narrow: b2Contact::Update's geometric half for ONE queued contact (lane per contact, not per env).
// Reads the two body transforms straight from the state words, runs SAT + clipping, matches impulses by feature id
// and writes the manifold + flags back.  Events are replayed later, in contact order, by finish_collide().
MRP_HD void narrow_item(const SimConst& K, const float* ct, uint32_t item) {
    const int64_t env = item / (uint32_t)kMaxC;
    const int k = (int)(item % (uint32_t)kMaxC);
    uint32_t* G = env_words(K, env);
    auto gw = [&](int w) -> uint32_t& { return G[w << kTileShift]; };
    auto gfl = [&](int w) { union { uint32_t u; float f; } c; c.u = G[w << kTileShift]; return c.f; };
    auto gsf = [&](int w, float v) { union { uint32_t u; float f; } c; c.f = v; G[w << kTileShift] = c.u; };
    const int cwk = K.w_con + k * MRP_CONTACT_WORDS;
    uint32_t m = gw(cwk);
    const int fa = m & 0xff, fb = (m >> 8) & 0xff;
    const int bA = (m >> 20) & 15, bB = (m >> 24) & 15;
    const int oldpc = (m >> 18) & 3;
    auto xf_of = [&](int b) {
        Xf x;
        if (b < K.nb) {
            const int w = K.w_body + kBodyWords * b;
            x.q.s = gfl(w + 6); x.q.c = gfl(w + 7);
            const V2 lc = b == 0 ? mk(K.blk_lcx, K.blk_lcy) : (b < K.nblk ? mk(K.blkx_lcx[b], K.blkx_lcy[b]) : mk(K.ag_lcx, K.ag_lcy));
            const V2 r = rmul(x.q, lc);
            x.p = K.hidden ? mk(gfl(w + 8), gfl(w + 9)) : mk(gfl(w + 0) - r.x, gfl(w + 1) - r.y);   // hidden: see Sim::load()
        } else {
            x.p = mk(ct[CT_WALLPOS + 2 * (b - K.nb)], ct[CT_WALLPOS + 2 * (b - K.nb) + 1]);
            x.q.s = 0.0f; x.q.c = 1.0f;
        }
        return x;
    };
    Manifold man;
    collide_polygons(&man, ct + CT_SHAPES + kShapeWords * (int)ct[CT_FIXSHAPE + fa], xf_of(bA),
                     ct + CT_SHAPES + kShapeWords * (int)ct[CT_FIXSHAPE + fb], xf_of(bB));
    const bool touching = man.pc > 0;
    if (touching) {
        const uint32_t okeys = oldpc ? gw(cwk + 1) : 0u;
        float nI[2] = {0.0f, 0.0f}, tI[2] = {0.0f, 0.0f};
        for (int i = 0; i < man.pc; ++i) {
            for (int j = 0; j < oldpc; ++j) {
                if (((okeys >> (16 * j)) & 0xffffu) == man.key[i]) {
                    nI[i] = gfl(cwk + 8 + 4 * j);
                    tI[i] = gfl(cwk + 9 + 4 * j);
                    break;
                }
            }
        }
        gw(cwk + 1) = man.key[0] | ((man.pc > 1 ? man.key[1] : 0u) << 16);
        gsf(cwk + 2, man.ln.x); gsf(cwk + 3, man.ln.y);
        gsf(cwk + 4, man.lp.x); gsf(cwk + 5, man.lp.y);
        for (int i = 0; i < man.pc; ++i) {
            gsf(cwk + 6 + 4 * i, man.pt[i].x);
            gsf(cwk + 7 + 4 * i, man.pt[i].y);
            gsf(cwk + 8 + 4 * i, nI[i]);
            gsf(cwk + 9 + 4 * i, tI[i]);
        }
    }
    m = (m & 0xfff0ffffu) | ((touching ? 1u : 0u) << 16) | (((uint32_t)man.type & 1u) << 17) | ((uint32_t)man.pc << 18);
    if (!touching) m &= ~(1u << 17);
    gw(cwk) = m;
}

}  // namespace mrp
