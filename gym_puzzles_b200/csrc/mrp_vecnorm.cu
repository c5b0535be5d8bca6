// mrp_vecnorm.cu — device-resident VecNormalize for the batched env.
//
// Replaces, for a batch that lives in HBM, what the reference's trainer wraps around its envs:
//   env = VecNormalize(env)                       reference train/train.py:82   (SB3 defaults: norm_obs, norm_reward,
//   env = VecNormalize.load(stats_path, env)      reference train/test.py:66-68  clip 10 / 10, gamma 0.99, epsilon 1e-8)
// i.e. running mean / variance of the observations and of the discounted returns (RunningMeanStd with Chan's
// parallel update, count initialised to 1e-4), obs -> clip((obs - mean) / sqrt(var + eps)), reward -> clip(reward /
// sqrt(var_ret + eps)), returns zeroed where done.  Stable-Baselines3 itself is a third-party dependency of the
// reference's training script (not in /root/reference); the arithmetic above is its published algorithm and is pinned
// by a numpy restatement in tests/test_vecnorm.py.
//
// Two HBM-bound passes per step over obs f32[N][O] (nothing here is a contraction):
//   k_vn_moments  reads obs (+ reward, done) once: per-column shifted sums  S1 = sum(x - m), S2 = sum((x - m)^2)  in
//                 float64 (m = the running mean, so the sums are well conditioned), returns <- returns*gamma + reward;
//                 warp-per-row-group, lane-per-column, block reduction in shared memory, one f64 atomicAdd per column
//                 and CTA into accum[2*(O+1)].  The accum vector is what ranks all-reduce (NCCL) in multi-GPU runs.
//   k_vn_merge    one CTA: folds the batch moments into the running statistics (Chan et al.), precomputes
//                 mean / inv-std as f32, clears accum.
//   k_vn_apply    elementwise normalise + clip, obs -> obs_out (may alias), reward -> reward_out, returns[done] = 0,
//                 optional terminal-observation rows of done envs.
// Algorithmic bytes per env-step: 4*O (moments read) + 8*O (apply read + write) + ~21 (reward, returns, done).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "mrp_vecnorm.h"

#ifndef MRP_HOST_EMU
#include <cuda_runtime.h>
#endif

#ifdef MRP_HOST_EMU
#define VN_HD inline
#else
#define VN_HD __host__ __device__ inline
#endif

namespace {

constexpr int kVnMaxObs = 96;              // columns (66 for v2 with 5 agents)
constexpr int kVnBlock = 256;
constexpr int kVnSlots = 8;               // shared-memory partial-sum copies per CTA (one per warp)
constexpr int kVnBlockM = 256;           // moments CTA (1024-thread CTAs at 2 per SM measured 2x slower: 32 registers per thread)

struct VnConst {
    int32_t N, O;
    int32_t norm_obs, norm_reward;
    float clip_obs, clip_reward;
    double epsilon, gamma;
    // running statistics (device): mean[O+1], var[O+1], count[2] (obs, returns); column O is the return
    double* mean;
    double* var;
    double* count;
    double* accum;      // [2*(O+1)] shifted batch sums, then [2] batch counts (obs rows, return rows)
    float* meanf;       // [O+1] f32 copies used by the apply pass
    float* istdf;       // [O+1] 1/sqrt(var + eps)
    double* returns;    // [N] discounted return per env
};

#ifndef MRP_HOST_EMU
// Flat float4 walk over obs[N*O]: `nthr` threads (a multiple of O / gcd(O, 4), so 4 * nthr is a multiple of O) each
// stride by nthr float4s, which keeps the four columns a thread touches fixed for the whole walk: running means and the
// eight f64 partial sums live in registers, four independent 16-byte loads are in flight per thread.
__global__ void __launch_bounds__(kVnBlockM) k_vn_moments(const __grid_constant__ VnConst V, const float* __restrict__ obs,
                                                          const float* __restrict__ rew, int64_t nthr) {
    // partial sums per warp (keeps the contention of the shared-memory f64 atomics low)
    __shared__ double redw[kVnSlots][2][kVnMaxObs + 1];
    const int O = V.O, slot = threadIdx.x / (kVnBlockM / kVnSlots);
    for (int i = threadIdx.x; i < kVnSlots * 2 * (kVnMaxObs + 1); i += kVnBlockM) (&redw[0][0][0])[i] = 0.0;
    __syncthreads();
    double (*red)[kVnMaxObs + 1] = redw[slot];
    const int64_t tid = (int64_t)blockIdx.x * kVnBlockM + threadIdx.x;
    if (obs && tid < nthr) {
        const int64_t total = (int64_t)V.N * O, n4 = total >> 2;
        int col[4];
        double m[4], s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            col[j] = (int)((4 * tid + j) % O);
            m[j] = (double)(float)V.mean[col[j]];
        }
        const float4* in4 = reinterpret_cast<const float4*>(obs);
        constexpr int U = 4;
        int64_t i = tid;
        for (; i + (U - 1) * nthr < n4; i += U * nthr) {
            float4 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = __ldg(in4 + i + u * nthr);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double d0 = (double)x[u].x - m[0], d1 = (double)x[u].y - m[1], d2 = (double)x[u].z - m[2], d3 = (double)x[u].w - m[3];
                s1[0] += d0; s2[0] += d0 * d0;
                s1[1] += d1; s2[1] += d1 * d1;
                s1[2] += d2; s2[2] += d2 * d2;
                s1[3] += d3; s2[3] += d3 * d3;
            }
        }
        for (; i < n4; i += nthr) {
            const float4 x = __ldg(in4 + i);
            const double d0 = (double)x.x - m[0], d1 = (double)x.y - m[1], d2 = (double)x.z - m[2], d3 = (double)x.w - m[3];
            s1[0] += d0; s2[0] += d0 * d0;
            s1[1] += d1; s2[1] += d1 * d1;
            s1[2] += d2; s2[2] += d2 * d2;
            s1[3] += d3; s2[3] += d3 * d3;
        }
        if (tid == 0)   // up to three trailing elements when N*O is not a multiple of 4
            for (int64_t e = n4 << 2; e < total; ++e) {
                const int c = (int)(e % O);
                const double d = (double)obs[e] - (double)(float)V.mean[c];
                atomicAdd(&red[0][c], d);
                atomicAdd(&red[1][c], d * d);
            }
#pragma unroll
        for (int j = 0; j < 4; ++j) { atomicAdd(&red[0][col[j]], s1[j]); atomicAdd(&red[1][col[j]], s2[j]); }
    }
    // discounted returns (column O): thread per env
    if (rew) {
        double r1 = 0.0, r2 = 0.0;
        const double mr = V.mean[O];
        for (int64_t e = tid; e < V.N; e += (int64_t)gridDim.x * kVnBlockM) {
            const double ret = V.returns[e] * V.gamma + (double)rew[e];
            V.returns[e] = ret;
            const double d = ret - mr;
            r1 += d;
            r2 += d * d;
        }
        for (int o = 16; o > 0; o >>= 1) { r1 += __shfl_xor_sync(0xffffffffu, r1, o); r2 += __shfl_xor_sync(0xffffffffu, r2, o); }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&red[0][O], r1); atomicAdd(&red[1][O], r2); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c <= O; c += kVnBlockM)
        if (c < O ? obs != nullptr : rew != nullptr) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < kVnSlots; ++w) { a += redw[w][0][c]; b += redw[w][1][c]; }
            atomicAdd(V.accum + c, a);
            atomicAdd(V.accum + (O + 1) + c, b);
        }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (obs) atomicAdd(V.accum + 2 * (O + 1), (double)V.N);
        if (rew) atomicAdd(V.accum + 2 * (O + 1) + 1, (double)V.N);
    }
}
#endif

// RunningMeanStd.update_from_moments (Chan et al.) for one column, from shifted sums around the running mean
VN_HD void vn_merge_col(double& mean, double& var, double count, double s1, double s2, double n, double shift) {
    const double bm_shift = s1 / n;                      // batch_mean - shift
    const double batch_var = s2 / n - bm_shift * bm_shift;
    const double delta = bm_shift + (shift - mean);      // batch_mean - mean
    const double tot = count + n;
    const double new_mean = mean + delta * n / tot;
    const double M2 = var * count + batch_var * n + delta * delta * count * n / tot;
    mean = new_mean;
    var = M2 / tot;
}

VN_HD void vn_merge(const VnConst& V, int col) {
    const int O = V.O;
    const int which = col < O ? 0 : 1;
    const double n = V.accum[2 * (O + 1) + which];
    if (n > 0.0) {
        double mean = V.mean[col], var = V.var[col];
        // obs columns were shifted by the f32 copy of the running mean (kept in a register per lane), the return by the f64 mean
        const double shift = col < O ? (double)(float)mean : mean;
        vn_merge_col(mean, var, V.count[which], V.accum[col], V.accum[(O + 1) + col], n, shift);
        V.mean[col] = mean;
        V.var[col] = var;
    }
    V.meanf[col] = (float)V.mean[col];
    V.istdf[col] = (float)(1.0 / sqrt(V.var[col] + V.epsilon));
}

#ifndef MRP_HOST_EMU
__global__ void k_vn_merge(const __grid_constant__ VnConst V) {
    const int O = V.O;
    for (int col = threadIdx.x; col <= O; col += blockDim.x) vn_merge(V, col);
    __syncthreads();
    if (threadIdx.x == 0) {
        V.count[0] += V.accum[2 * (O + 1)];
        V.count[1] += V.accum[2 * (O + 1) + 1];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * (O + 1) + 2; i += blockDim.x) V.accum[i] = 0.0;
}

__device__ __forceinline__ float vn_norm(float x, float m, float is, float clip) {
    const float y = (x - m) * is;
    return fminf(fmaxf(y, -clip), clip);
}

// obs_out may alias obs (in-place normalisation, include/mrp_vecnorm.h): neither is __restrict__ nor read through the
// non-coherent path; every thread reads its elements before it writes them
__global__ void __launch_bounds__(kVnBlock) k_vn_apply(const __grid_constant__ VnConst V, const float* obs,
                                                        const float* rew, const uint8_t* __restrict__ done,
                                                        float* obs_out, float* rew_out, float* term_obs) {
    __shared__ float sm[kVnMaxObs + 1], si[kVnMaxObs + 1];
    const int O = V.O;
    for (int c = threadIdx.x; c <= O; c += kVnBlock) { sm[c] = V.meanf[c]; si[c] = V.istdf[c]; }
    __syncthreads();
    const int64_t total = (int64_t)V.N * O, stride = (int64_t)gridDim.x * kVnBlock;
    const int64_t tid = (int64_t)blockIdx.x * kVnBlock + threadIdx.x;
    if (obs_out) {
        // flat float4 walk (obs is contiguous, so this holds for any O); columns advance with wrap-around
        const int64_t n4 = total >> 2;
        const float4* in4 = reinterpret_cast<const float4*>(obs);
        float4* out4 = reinterpret_cast<float4*>(obs_out);
        for (int64_t i = tid; i < n4; i += stride) {
            float4 v = in4[i];
            if (V.norm_obs) {
                int c0 = (int)((i * 4) % O);
                int c1 = c0 + 1 == O ? 0 : c0 + 1;
                int c2 = c1 + 1 == O ? 0 : c1 + 1;
                int c3 = c2 + 1 == O ? 0 : c2 + 1;
                v.x = vn_norm(v.x, sm[c0], si[c0], V.clip_obs);
                v.y = vn_norm(v.y, sm[c1], si[c1], V.clip_obs);
                v.z = vn_norm(v.z, sm[c2], si[c2], V.clip_obs);
                v.w = vn_norm(v.w, sm[c3], si[c3], V.clip_obs);
            }
            out4[i] = v;
        }
        if (tid == 0)
            for (int64_t e = n4 << 2; e < total; ++e) {
                const float x = obs[e];
                obs_out[e] = V.norm_obs ? vn_norm(x, sm[(int)(e % O)], si[(int)(e % O)], V.clip_obs) : x;
            }
    }
    for (int64_t e = tid; e < V.N; e += stride) {
        const bool d = done && done[e];
        if (rew && rew_out) {
            const float r = rew[e];
            rew_out[e] = V.norm_reward ? fminf(fmaxf(r * si[O], -V.clip_reward), V.clip_reward) : r;
        }
        if (d) {
            V.returns[e] = 0.0;
            if (term_obs && V.norm_obs) {
                float* row = term_obs + e * O;
                for (int c = 0; c < O; ++c) row[c] = vn_norm(row[c], sm[c], si[c], V.clip_obs);
            }
        }
    }
}
#endif

}  // namespace

struct mrp_vecnorm {
    VnConst V;
    int device;
    int num_sms;   // multiprocessors of the device (148 on B200): the persistent grids are sized from it
    int training;
    int64_t launches;
};

static thread_local char g_vn_err[256] = "";
static int vn_fail(int code, const char* msg) {
    snprintf(g_vn_err, sizeof(g_vn_err), "%s", msg);
    return code;
}

#ifdef MRP_HOST_EMU
#define VN_ALLOC(ptr, bytes) ((*(void**)&(ptr) = calloc(1, (bytes))) ? 0 : -1)
#define VN_FREE(ptr) free(ptr)
#define VN_H2D(dst, src, bytes) memcpy((dst), (src), (bytes))
#define VN_D2H(dst, src, bytes) memcpy((dst), (src), (bytes))
#else
#define VN_ALLOC(ptr, bytes) (cudaMalloc((void**)&(ptr), (bytes)) == cudaSuccess ? (cudaMemset((ptr), 0, (bytes)), 0) : -1)
#define VN_FREE(ptr) cudaFree(ptr)
#define VN_H2D(dst, src, bytes) cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice)
#define VN_D2H(dst, src, bytes) cudaMemcpy((dst), (src), (bytes), cudaMemcpyDeviceToHost)
#endif

extern "C" {

const char* mrp_vecnorm_last_error(void) { return g_vn_err; }

int mrp_vecnorm_destroy(mrp_vecnorm* vn) {
    if (!vn) return 0;
#ifndef MRP_HOST_EMU
    cudaSetDevice(vn->device);
#endif
    VN_FREE(vn->V.mean); VN_FREE(vn->V.var); VN_FREE(vn->V.count); VN_FREE(vn->V.accum);
    VN_FREE(vn->V.meanf); VN_FREE(vn->V.istdf); VN_FREE(vn->V.returns);
    delete vn;
    return 0;
}

int mrp_vecnorm_set_stats(mrp_vecnorm* vn, const double* stats_host);

int mrp_vecnorm_create(const mrp_vecnorm_config* cfg, mrp_vecnorm** out) {
    if (!cfg || !out) return vn_fail(-1, "mrp_vecnorm_create: null argument");
    *out = nullptr;
    if (cfg->num_envs <= 0 || cfg->obs_dim <= 0 || cfg->obs_dim > kVnMaxObs) return vn_fail(-2, "mrp_vecnorm_create: bad num_envs / obs_dim (<= 96)");
#ifndef MRP_HOST_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev)
        return vn_fail(-3, "mrp_vecnorm_create: no such CUDA device; this library has no CPU path");
    cudaSetDevice(cfg->device);
#endif
    mrp_vecnorm* vn = new (std::nothrow) mrp_vecnorm();
    if (!vn) return vn_fail(-5, "mrp_vecnorm_create: out of host memory");
    memset(vn, 0, sizeof(*vn));
    VnConst& V = vn->V;
    V.N = cfg->num_envs; V.O = cfg->obs_dim;
    V.norm_obs = cfg->norm_obs; V.norm_reward = cfg->norm_reward;
    V.clip_obs = (float)cfg->clip_obs; V.clip_reward = (float)cfg->clip_reward; V.epsilon = cfg->epsilon; V.gamma = cfg->gamma;
    vn->device = cfg->device;
    vn->num_sms = 148;
#ifndef MRP_HOST_EMU
    if (cudaDeviceGetAttribute(&vn->num_sms, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess || vn->num_sms < 1) vn->num_sms = 148;
#endif
    vn->training = cfg->training;
    const size_t C = (size_t)V.O + 1;
    int rc = VN_ALLOC(V.mean, sizeof(double) * C) | VN_ALLOC(V.var, sizeof(double) * C) | VN_ALLOC(V.count, sizeof(double) * 2) |
             VN_ALLOC(V.accum, sizeof(double) * (2 * C + 2)) | VN_ALLOC(V.meanf, sizeof(float) * C) | VN_ALLOC(V.istdf, sizeof(float) * C) |
             VN_ALLOC(V.returns, sizeof(double) * (size_t)V.N);
    if (rc) { mrp_vecnorm_destroy(vn); return vn_fail(-7, "mrp_vecnorm_create: device allocation failed"); }
    // RunningMeanStd(epsilon=1e-4): mean 0, var 1, count 1e-4
    double* init = (double*)calloc(2 * C + 2, sizeof(double));
    for (size_t i = 0; i < C; ++i) init[C + i] = 1.0;
    init[2 * C] = 1e-4; init[2 * C + 1] = 1e-4;
    mrp_vecnorm_set_stats(vn, init);
    free(init);
    *out = vn;
    return 0;
}

int mrp_vecnorm_set_training(mrp_vecnorm* vn, int32_t training) {
    if (!vn) return vn_fail(-1, "mrp_vecnorm_set_training: null handle");
    vn->training = training ? 1 : 0;
    return 0;
}

// stats layout (host, f64): mean[O+1], var[O+1], count_obs, count_ret  — column O is the discounted return
int mrp_vecnorm_get_stats(mrp_vecnorm* vn, double* stats_host) {
    if (!vn || !stats_host) return vn_fail(-1, "mrp_vecnorm_get_stats: null argument");
    const size_t C = (size_t)vn->V.O + 1;
#ifndef MRP_HOST_EMU
    cudaSetDevice(vn->device);
    cudaDeviceSynchronize();
#endif
    VN_D2H(stats_host, vn->V.mean, sizeof(double) * C);
    VN_D2H(stats_host + C, vn->V.var, sizeof(double) * C);
    VN_D2H(stats_host + 2 * C, vn->V.count, sizeof(double) * 2);
    return 0;
}

int mrp_vecnorm_set_stats(mrp_vecnorm* vn, const double* stats_host) {
    if (!vn || !stats_host) return vn_fail(-1, "mrp_vecnorm_set_stats: null argument");
    const VnConst& V = vn->V;
    const size_t C = (size_t)V.O + 1;
#ifndef MRP_HOST_EMU
    cudaSetDevice(vn->device);
    cudaDeviceSynchronize();
#endif
    VN_H2D(V.mean, stats_host, sizeof(double) * C);
    VN_H2D(V.var, stats_host + C, sizeof(double) * C);
    VN_H2D(V.count, stats_host + 2 * C, sizeof(double) * 2);
    float* mf = (float*)malloc(sizeof(float) * 2 * C);
    for (size_t i = 0; i < C; ++i) {
        mf[i] = (float)stats_host[i];
        mf[C + i] = (float)(1.0 / sqrt(stats_host[C + i] + V.epsilon));
    }
    VN_H2D(V.meanf, mf, sizeof(float) * C);
    VN_H2D(V.istdf, mf + C, sizeof(float) * C);
    free(mf);
    return 0;
}

// the vector ranks sum with one all-reduce between mrp_vecnorm_moments and mrp_vecnorm_apply: f64[2*(O+1)+2]
int mrp_vecnorm_accum(mrp_vecnorm* vn, double** accum_dev, int32_t* count) {
    if (!vn || !accum_dev || !count) return vn_fail(-1, "mrp_vecnorm_accum: null argument");
    *accum_dev = vn->V.accum;
    *count = 2 * (vn->V.O + 1) + 2;
    return 0;
}

// pass 1 (training only): batch moments of obs (and of the discounted returns when reward_dev != NULL; pass NULL
// after a reset, where SB3 updates obs_rms only)
int mrp_vecnorm_moments(mrp_vecnorm* vn, const float* obs_dev, const float* reward_dev, void* stream) {
    if (!vn || !obs_dev) return vn_fail(-1, "mrp_vecnorm_moments: null argument");
    if (!vn->training) return 0;
    const VnConst& V = vn->V;
#ifndef MRP_HOST_EMU
    cudaSetDevice(vn->device);
    // persistent grid: 8 CTAs of 256 threads per SM (fewer for small batches); the obs walk uses the largest thread
    // count that is a multiple of O / gcd(O, 4)
    const int64_t n4 = ((int64_t)V.N * V.O) >> 2;
    int64_t grid = (n4 / 4 + kVnBlockM - 1) / kVnBlockM;
    if (grid > (int64_t)vn->num_sms * 8) grid = (int64_t)vn->num_sms * 8;
    if (grid < 1) grid = 1;
    int g = 4, o = V.O;
    while (o) { int t = g % o; g = o; o = t; }   // gcd(4, O)
    const int64_t L = V.O / g;
    int64_t nthr = grid * kVnBlockM / L * L;
    if (nthr < L) { grid = (L + kVnBlockM - 1) / kVnBlockM; nthr = L; }
    k_vn_moments<<<(unsigned)grid, kVnBlockM, 0, (cudaStream_t)stream>>>(V, V.norm_obs ? obs_dev : nullptr, V.norm_reward ? reward_dev : nullptr, nthr);
    vn->launches += 1;
    if (cudaGetLastError() != cudaSuccess) return vn_fail(-10, "mrp_vecnorm_moments: launch failed");
#else
    (void)stream;
    const int O = V.O;
    for (int c = 0; c < O && V.norm_obs; ++c) {
        const float m = (float)V.mean[c];
        double s1 = 0.0, s2 = 0.0;
        for (int64_t r = 0; r < V.N; ++r) { const double d = (double)obs_dev[r * O + c] - (double)m; s1 += d; s2 += d * d; }
        V.accum[c] += s1; V.accum[(O + 1) + c] += s2;
    }
    if (V.norm_obs) V.accum[2 * (O + 1)] += (double)V.N;
    if (reward_dev && V.norm_reward) {
        double s1 = 0.0, s2 = 0.0;
        for (int64_t e = 0; e < V.N; ++e) {
            const double ret = V.returns[e] * V.gamma + (double)reward_dev[e];
            V.returns[e] = ret;
            const double d = ret - V.mean[O];
            s1 += d; s2 += d * d;
        }
        V.accum[O] += s1; V.accum[(O + 1) + O] += s2;
        V.accum[2 * (O + 1) + 1] += (double)V.N;
    }
#endif
    return 0;
}

// pass 2: fold the (possibly all-reduced) batch moments into the running statistics, then normalise.
// obs_out_dev may alias obs_dev; reward_dev / reward_out_dev / done_dev / terminal_obs_dev may be NULL.
int mrp_vecnorm_apply(mrp_vecnorm* vn, const float* obs_dev, const float* reward_dev, const uint8_t* done_dev, float* obs_out_dev,
                      float* reward_out_dev, float* terminal_obs_dev, void* stream) {
    if (!vn || !obs_dev) return vn_fail(-1, "mrp_vecnorm_apply: null argument");
    const VnConst& V = vn->V;
#ifndef MRP_HOST_EMU
    cudaSetDevice(vn->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (vn->training) { k_vn_merge<<<1, 128, 0, st>>>(V); vn->launches += 1; }
    const int64_t work = ((int64_t)V.N * V.O + 3) / 4;
    int64_t grid = (work + kVnBlock - 1) / kVnBlock;
    if (grid > (int64_t)vn->num_sms * 16) grid = (int64_t)vn->num_sms * 16;
    k_vn_apply<<<(unsigned)grid, kVnBlock, 0, st>>>(V, obs_dev, reward_dev, done_dev, obs_out_dev, reward_out_dev, terminal_obs_dev);
    vn->launches += 1;
    if (cudaGetLastError() != cudaSuccess) return vn_fail(-10, "mrp_vecnorm_apply: launch failed");
#else
    (void)stream;
    const int O = V.O;
    if (vn->training) {
        for (int c = 0; c <= O; ++c) vn_merge(V, c);
        V.count[0] += V.accum[2 * (O + 1)];
        V.count[1] += V.accum[2 * (O + 1) + 1];
        for (int i = 0; i < 2 * (O + 1) + 2; ++i) V.accum[i] = 0.0;
    }
    auto norm = [&](float x, int c) {
        const float y = (x - V.meanf[c]) * V.istdf[c];
        return fminf(fmaxf(y, -V.clip_obs), V.clip_obs);
    };
    if (obs_out_dev)
        for (int64_t i = 0; i < (int64_t)V.N * O; ++i) obs_out_dev[i] = V.norm_obs ? norm(obs_dev[i], (int)(i % O)) : obs_dev[i];
    for (int64_t e = 0; e < V.N; ++e) {
        if (reward_dev && reward_out_dev) {
            const float r = reward_dev[e];
            reward_out_dev[e] = V.norm_reward ? fminf(fmaxf(r * V.istdf[O], -V.clip_reward), V.clip_reward) : r;
        }
        if (done_dev && done_dev[e]) {
            V.returns[e] = 0.0;
            if (terminal_obs_dev && V.norm_obs)
                for (int c = 0; c < O; ++c) terminal_obs_dev[e * O + c] = norm(terminal_obs_dev[e * O + c], c);
        }
    }
#endif
    return 0;
}

// returns <- 0 (VecNormalize.reset, vec_normalize.py: self.returns = np.zeros(num_envs))
int mrp_vecnorm_reset_returns(mrp_vecnorm* vn, void* stream) {
    if (!vn) return vn_fail(-1, "mrp_vecnorm_reset_returns: null handle");
#ifndef MRP_HOST_EMU
    cudaSetDevice(vn->device);
    cudaMemsetAsync(vn->V.returns, 0, sizeof(double) * (size_t)vn->V.N, (cudaStream_t)stream);
#else
    (void)stream;
    memset(vn->V.returns, 0, sizeof(double) * (size_t)vn->V.N);
#endif
    return 0;
}

int64_t mrp_vecnorm_launch_count(mrp_vecnorm* vn) { return vn ? vn->launches : 0; }

}  // extern "C"
