/*
 * mrp_vecnorm.h — C-ABI of the device-resident VecNormalize that follows the batched env.
 *
 * Replaces, for observation / reward tensors that live in HBM, the wrapper the reference's trainer puts around its
 * envs:  env = VecNormalize(env)  (reference train/train.py:82; evaluation reload train/test.py:66-68), i.e.
 * Stable-Baselines3's VecNormalize with its defaults (norm_obs, norm_reward, clip_obs = clip_reward = 10,
 * gamma = 0.99, epsilon = 1e-8, RunningMeanStd count initialised to 1e-4).  SB3 is a third-party dependency of the
 * training script and is not part of /root/reference; tests/test_vecnorm.py pins the arithmetic with a numpy
 * restatement of its published algorithm.
 *
 * Conventions as in mrp_b200.h: 0 on success, negative on error (mrp_vecnorm_last_error()), stream-ordered device
 * calls, no torch types, no CPU implementation behind it.
 */
#ifndef MRP_VECNORM_H
#define MRP_VECNORM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mrp_vecnorm mrp_vecnorm;

typedef struct mrp_vecnorm_config {
    int32_t num_envs, obs_dim, device;
    int32_t norm_obs, norm_reward, training; /* SB3 defaults: 1, 1, 1 */
    double clip_obs, clip_reward;            /* 10, 10 */
    double gamma, epsilon;                   /* 0.99, 1e-8 (Python floats in SB3) */
} mrp_vecnorm_config;

const char* mrp_vecnorm_last_error(void);
int mrp_vecnorm_create(const mrp_vecnorm_config* cfg, mrp_vecnorm** out);
int mrp_vecnorm_destroy(mrp_vecnorm* vn);
/* VecNormalize.training (train/test.py:67 sets it False for evaluation): statistics frozen when 0 */
int mrp_vecnorm_set_training(mrp_vecnorm* vn, int32_t training);

/* pass 1 (no-op unless training): shifted batch moments of obs f32[num_envs][obs_dim] and, when reward_dev != NULL,
 * of the discounted returns (returns <- returns*gamma + reward) into the accumulator vector. */
int mrp_vecnorm_moments(mrp_vecnorm* vn, const float* obs_dev, const float* reward_dev, void* stream);
/* the accumulator f64[count] on the device: in multi-GPU runs every rank sums it with ONE all-reduce (NCCL) between
 * the two passes, so all ranks keep identical running statistics */
int mrp_vecnorm_accum(mrp_vecnorm* vn, double** accum_dev, int32_t* count);
/* pass 2: fold the batch moments into the running statistics (Chan's parallel update), then
 * obs_out = clip((obs - mean) / sqrt(var + eps)), reward_out = clip(reward / sqrt(var_ret + eps)),
 * returns[done] = 0, terminal-observation rows of done envs normalised in place.
 * obs_out_dev may alias obs_dev and reward_out_dev may alias reward_dev (in place); reward_dev, reward_out_dev, done_dev, terminal_obs_dev may be NULL. */
int mrp_vecnorm_apply(mrp_vecnorm* vn, const float* obs_dev, const float* reward_dev, const uint8_t* done_dev,
                      float* obs_out_dev, float* reward_out_dev, float* terminal_obs_dev, void* stream);
/* VecNormalize.reset(): returns <- 0 */
int mrp_vecnorm_reset_returns(mrp_vecnorm* vn, void* stream);

/* save / load (VecNormalize.save / .load keep exactly these numbers in saved_env.pkl, train/train.py:149):
 * f64 host vector  mean[obs_dim+1], var[obs_dim+1], count_obs, count_ret  — column obs_dim is the return */
int mrp_vecnorm_get_stats(mrp_vecnorm* vn, double* stats_host);
int mrp_vecnorm_set_stats(mrp_vecnorm* vn, const double* stats_host);

int64_t mrp_vecnorm_launch_count(mrp_vecnorm* vn);

#ifdef __cplusplus
}
#endif
#endif /* MRP_VECNORM_H */
