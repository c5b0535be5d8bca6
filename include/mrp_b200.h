/*
 * mrp_b200.h — C-ABI of the B200-native batched MultiRobotPuzzle simulator.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own
 * for this path: gym_puzzles reaches Box2D through pybox2d's SWIG objects, one
 * b2World per Python env.  Each entry point below names the reference call(s) it
 * replaces (paths relative to /root/reference, mrp00 =
 * gym_puzzles/envs/multi_robot_puzzle_00.py, mrp02 = .../multi_robot_puzzle_02.py).
 *
 * Conventions: every function returns 0 on success and a negative code on error;
 * mrp_last_error() returns a thread-local message.  No exceptions cross the ABI, no
 * torch types appear in it.  A handle is bound to one CUDA device and is not
 * thread-safe; different handles are independent.  All `*_dev` pointers are device
 * pointers on the handle's device; `stream` is a cudaStream_t passed as void*
 * (NULL = legacy default stream).  Stream-ordered calls never synchronise the host.
 * There is no CPU implementation behind this ABI: without a CUDA device
 * mrp_create() fails.
 *
 * Device memory of a handle (MultiRobotPuzzleHeavy-v0, per env): state 2.2 KB, solver task pool 4.9 KB (a worst-case
 * reservation, touched only for touching contacts), queues 0.5 KB, action / observation / reward rows 0.23 KB; from
 * 131,072 envs also the spare episodes (next episodes computed ahead of time, so that an
 * auto-reset is a copy): a second state buffer and observation buffer (2.4 KB) plus the queues of the refill pass for an
 * eighth of the batch (0.7 KB).  1,048,576 envs: ~11 GB.  Steps of batches below 32,768 envs are replayed from a CUDA
 * graph; mrp_step_host pipelines large batches (front-half waves, chunked result copies).  None of this changes
 * results; the environment variables that switch it are listed in INTEGRATION.md §4.
 */
#ifndef MRP_B200_H
#define MRP_B200_H

#include <stdint.h>

#include "mrp_state.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mrp_handle mrp_handle;

/* replaces: gym.make(id) -> MultiRobotPuzzle.__init__ (mrp00:152-209),
 * MultiRobotPuzzle2.__init__(frameskip=1, num_agents=2) (mrp02:139-197) and the
 * registration kwargs max_episode_steps (gym_puzzles/__init__.py:3-29), for a batch. */
typedef struct mrp_config {
    int32_t variant;      /* MRP_VARIANT_* */
    int32_t n_agents;     /* <=0: registered default (2; 5 for Heavy-v0); v2 ctor kw num_agents */
    int32_t num_envs;     /* envs held by this handle (this GPU's shard) */
    int32_t device;       /* CUDA device ordinal */
    uint64_t seed;        /* Philox key (spawn / hidden reset action / synthetic action streams) */
    uint64_t env_id_base; /* global id of env 0 of this shard: RNG is keyed by global id, so results
                             do not depend on how envs are sharded over GPUs */
    int32_t auto_reset;   /* 1: gym-0.21 vector semantics — a done env is reset inside step() */
    int32_t max_episode_steps; /* <=0: registered TimeLimit (2000 / 3000 Heavy-v0) */
} mrp_config;

/* library-owned device buffers, valid for the handle's lifetime; Python wraps them zero-copy */
typedef struct mrp_buffers {
    float* action_dev;   /* f32[num_envs][act_dim]   read by mrp_step when actions_dev == NULL */
    float* obs_dev;      /* f32[num_envs][obs_dim]   mrp00:441-472 / mrp02:491-532 */
    float* reward_dev;   /* f32[num_envs]            mrp00:474-519 / mrp02:534-582 */
    uint8_t* done_dev;   /* u8[num_envs]             env done OR TimeLimit */
    uint8_t* trunc_dev;  /* u8[num_envs]             info['TimeLimit.truncated'] */
    double* stats_dev;   /* f64[MRP_N_STATS]         episode statistics (NCCL-allreduce this) */
    int32_t num_envs, obs_dim, act_dim, reserved;
} mrp_buffers;

/* stats_dev slots.  0-7: episode statistics.  8: envs whose body state became NaN / inf and were reset by force
 * (SURVEY.md §5 / §8b "NaN guard counter"; such a step reports done = trunc = 1, reward 0).  9-14: workload counters
 * of the steps since the last reset of the statistics, for the FLOP roofline of SURVEY.md §8d
 * (F = F_fix + 700 P + 180 (81 m1 + 160 m2) + 64 sum(position points) + F_toi): */
enum {
    MRP_STAT_EPISODES = 0, MRP_STAT_DONE_BY_ENV = 1, MRP_STAT_TRUNCATED = 2, MRP_STAT_SUM_RETURN = 3,
    MRP_STAT_SUM_RETURN_SQ = 4, MRP_STAT_SUM_LENGTH = 5, MRP_STAT_ENV_STEPS = 6, MRP_STAT_OVERFLOW = 7,
    MRP_STAT_NAN_RESETS = 8,
    MRP_STAT_PAIRS = 9,        /* P: narrowphase (SAT + clip) evaluations actually run (culled pairs not counted) */
    MRP_STAT_M1 = 10,          /* touching manifolds handed to the solver with one point ... */
    MRP_STAT_M2 = 11,          /* ... and with two points (block solver): Box2D runs 180 sweeps over each */
    MRP_STAT_VEL_FLOPS = 12,   /* velocity-sweep flops actually executed (81 / 160 per contact and sweep; early exit) */
    MRP_STAT_POS_POINTS = 13,  /* manifold-point corrections executed by the position solver (64 flops each) */
    MRP_STAT_TOI_CALLS = 14,   /* b2TimeOfImpact evaluations (culled sweeps not counted) */
    MRP_STAT_RESERVED = 15,
    MRP_N_STATS = 16
};

/* replaces: set_reward_params (mrp00:231-239, mrp02:216-225), update_goal (mrp02:232-233),
 * update_params (mrp02:227-230; decay_pow = decay**(-timestep), default 1 — SURVEY.md C.1) */
typedef struct mrp_params {
    double agentDelta, agentDistance, blockDelta, blockDistance;
    double puzzleComp, outOfBounds, blkOutOfBounds;
    double scaled_epsilon;
    double decay_pow;
} mrp_params;

const char* mrp_last_error(void);
/* "cuda-sm_100a" for the product library */
const char* mrp_backend(void);

int mrp_create(const mrp_config* cfg, mrp_handle** out);
int mrp_destroy(mrp_handle* h);
int mrp_get_layout(mrp_handle* h, mrp_layout* out);
int mrp_get_buffers(mrp_handle* h, mrp_buffers* out);

/* replaces: env.reset() (mrp00:392-411, mrp02:421-442): respawn + ONE hidden physics step
 * with a sampled action; writes obs_dev rows of the reset envs.
 * mask_dev: u8[num_envs] device mask or NULL (= all envs). */
int mrp_reset(mrp_handle* h, const uint8_t* mask_dev, void* stream);

/* replaces: env.step(action) (mrp00:413-521, mrp02:444-584) incl. world.Step(1/50, 180, 60)
 * (mrp00:428, mrp02:478), the TimeLimit wrapper and (auto_reset) the vector-env reset.
 * actions_dev: f32[num_envs][act_dim] or NULL (= the owned action buffer). */
int mrp_step(mrp_handle* h, const float* actions_dev, void* stream);

/* Host-buffer convenience forms (what a non-CUDA-aware caller of the reference would use):
 * H2D copy of actions, step, D2H copies of the results, then a stream synchronise.
 * Any output pointer may be NULL.  Pinned (page-locked, device-visible) buffers make the copies overlap the kernels;
 * with a pinned obs buffer the rows of a chunk of envs leave before that chunk's TOI-event / auto-reset passes and the
 * few rows those passes rewrite are stored into the buffer by a kernel.  Pageable buffers work, more slowly.
 * Ordering: mrp_step_host orders its work after what the caller queued on the legacy default stream only; a caller
 * that mixes it with mrp_step / mrp_reset on another stream synchronises that stream first.  mrp_reset_host,
 * mrp_get_state, mrp_set_state and mrp_get_stats synchronise the device on entry. */
int mrp_step_host(mrp_handle* h, const float* actions_host, float* obs_host, float* reward_host,
                  uint8_t* done_host, uint8_t* trunc_host);
int mrp_reset_host(mrp_handle* h, const uint8_t* mask_host, float* obs_host);

/* synthetic benchmark actions: U(-1,1) f32, Philox stream ACTION,
 * counter (global env id, step_index); dst_dev NULL = owned action buffer. */
int mrp_sample_actions(mrp_handle* h, uint64_t step_index, float* dst_dev, void* stream);

/* canonical state records (include/mrp_state.h), host memory; synchronous.
 * No reference equivalent ("Checkpoint / resume: none"). */
int mrp_get_state(mrp_handle* h, int32_t env_begin, int32_t env_count, uint32_t* words_host);
int mrp_set_state(mrp_handle* h, int32_t env_begin, int32_t env_count, const uint32_t* words_host);

/* "Next" rows of SURVEY.md §8f — what the reference's callers (train/train.py:63-82: Monitor, DummyVecEnv,
 * VecNormalize) need from a batched env that resets inside step():
 * the last observation of a finished episode (SB3 infos[i]["terminal_observation"]) and its return / length
 * (Monitor's info["episode"] = {"r", "l"}), kept in library-owned buffers before the auto-reset overwrites them.
 * Rows are valid for envs whose done flag is set by the same step.  Off (no cost) until enabled. */
typedef struct mrp_terminal_buffers {
    float* terminal_obs_dev;     /* f32[num_envs][obs_dim] */
    float* episode_return_dev;   /* f32[num_envs] */
    int32_t* episode_length_dev; /* i32[num_envs] */
} mrp_terminal_buffers;
int mrp_enable_terminal_info(mrp_handle* h, mrp_terminal_buffers* out);

/* Per-env curriculum: update_goal(epoch, nb_epochs) (mrp02:232-233) and update_params(timestep, decay)
 * (mrp02:227-230) as vectors, one value per env, so that envs of one batch can sit at different curriculum stages.
 * Returns library-owned device arrays f64[num_envs] initialised from the scalar mrp_params; the caller writes them
 * (stream-ordered before mrp_step).  v2 family only (v0 uses the fixed EPSILON = 25 px, mrp00:54). */
int mrp_enable_curriculum(mrp_handle* h, double** scaled_epsilon_dev, double** decay_pow_dev);

/* Alternative observation head for the holonomic family: the normalised observation of the reference's experimental
 * MultiRobotPuzzle-v3 (gym_puzzles/envs/core.py:289-350: _get_norm_pose, _get_obs) computed from the current state —
 * out_dev: f32[num_envs][4 * n_agents + 19] = per robot {bx-ax, by-ay, rot mod 2pi, contact}, block {gx-bx, gy-by, -rot},
 * 8 vertices {(x-w/2)/(w/2), (y-h/2)/(w/2)}.  Stream-ordered; call after mrp_step / mrp_reset. */
int mrp_obs_v3(mrp_handle* h, float* out_dev, void* stream);

int mrp_set_params(mrp_handle* h, const mrp_params* p);
int mrp_get_params(mrp_handle* h, mrp_params* p);

/* copies stats_dev to host (synchronous); reset_after != 0 zeroes the device counters */
int mrp_get_stats(mrp_handle* h, double* out_host, int32_t reset_after);

/* Device timing of the step's kernels alone (CUDA events recorded around them on the caller's stream), for
 * bench.py's roofline: enable, run steps, then read the accumulated milliseconds / step count
 * (mrp_get_timing synchronises on the recorded events). */
int mrp_set_timing(mrp_handle* h, int32_t enable);
int mrp_get_timing(mrp_handle* h, double* total_ms, int64_t* count, int32_t reset_after);
/* accumulated milliseconds of the five phase kernels of mrp_step, in launch order:
 * k_pre, k_solve_vel, k_solve_pos, k_post, k_post_events */
int mrp_get_phase_timing(mrp_handle* h, double* ms5, int32_t reset_after);

/* number of kernels this handle has launched so far (bench.py "gpu_launches") */
int64_t mrp_launch_count(mrp_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* MRP_B200_H */
