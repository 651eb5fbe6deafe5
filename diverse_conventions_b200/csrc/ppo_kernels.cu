// ppo_kernels.cu — returns / GAE and advantage normalisation directly over the seat-major rollout
// buffer that ocb_rollout_policy fills.
//
// Reference: SharedReplayBuffer.compute_returns (train/MAPPO/utils/shared_buffer.py:248-304),
// ValueNorm.denormalize (train/MAPPO/utils/valuenorm.py:76-87) and the advantage normalisation at the
// top of R_MAPPO.train (train/MAPPO/r_mappo.py:174-182).  The reference walks the T steps in a
// Python loop of ~10 torch ops on [N,2,1] tensors; here one thread owns one (seat, world) column
// and walks it backwards with the loads of the next steps already in flight, so the pass is one
// read of value_preds / rewards / dones and one write of returns / advantages (HBM-bound,
// 16 + 4/P bytes per agent-step).
//
// Arithmetic follows the reference operation by operation in fp32 with explicit round-to-nearest
// multiplies and adds (no FMA contraction), so returns and un-normalised advantages are
// bit-identical to torch's CPU result; the mean / std reduction runs in fp64.
#include <cuda_runtime.h>
#include <stdint.h>

#include "api_common.h"
#include "ocb.h"

using namespace ocb;

namespace {

constexpr int kGaeThreads = 64;   // small blocks: P*N columns are few (32,768 at config 4), spread them over all SMs
constexpr int kGaeUnroll = 16;    // steps per batch; two batches of loads are in flight per thread (software pipeline)

struct GaeParams {
    const float* value_preds;  // [T+1][R]
    const int32_t* rewards;    // [T][R]
    const int32_t* done;       // [T][N]
    float* returns;            // [T+1][R]
    float* advantages;         // [T][R] or nullptr
    double* stats;             // [3] sum, sum of squares, count (or nullptr)
    int T, R, N;
    float gamma, gl, vn_mean, vn_std;
    int use_gae;
};

__device__ __forceinline__ float denorm(float v, float std, float mean) { return __fadd_rn(__fmul_rn(v, std), mean); }

struct GaeBatch {
    float v[kGaeUnroll];
    int32_t r[kGaeUnroll], d[kGaeUnroll];
};
// loads of steps t1-1 .. t1-kGaeUnroll (clamped at 0; clamped entries are never consumed)
__device__ __forceinline__ void gae_load(GaeBatch& b, const float* vp, const int32_t* rw, const int32_t* dn, int t1, size_t R, size_t N) {
#pragma unroll
    for (int k = 0; k < kGaeUnroll; ++k) {
        const int t = t1 - 1 - k;
        const size_t tt = t >= 0 ? (size_t)t : 0;
        b.v[k] = __ldcs(vp + tt * R);
        b.r[k] = __ldcs(rw + tt * R);
        b.d[k] = __ldg(dn + tt * N);
    }
}

__global__ void __launch_bounds__(kGaeThreads) gae_kernel(const GaeParams p) {
    const int col = blockIdx.x * kGaeThreads + threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    if (col < p.R) {
        const int n = col % p.N;
        const size_t R = (size_t)p.R, N = (size_t)p.N;
        const float* vp = p.value_preds + col;
        const int32_t* rw = p.rewards + col;
        const int32_t* dn = p.done + n;
        float* ret = p.returns + col;
        float* adv = p.advantages ? p.advantages + col : nullptr;
        GaeBatch cur, nxt;
        gae_load(cur, vp, rw, dn, p.T, R, N);
        const float v_next = vp[(size_t)p.T * R];
        float gae = 0.0f;
        float ret_next = v_next;  // discounted-sum branch: returns[T] = next_value (shared_buffer.py:297)
        if (!p.use_gae) ret[(size_t)p.T * R] = v_next;  // the GAE branch never writes returns[T] (shared_buffer.py:277-287)
        float dn_next = denorm(v_next, p.vn_std, p.vn_mean);
        for (int t1 = p.T; t1 > 0; t1 -= kGaeUnroll) {
            if (t1 > kGaeUnroll) gae_load(nxt, vp, rw, dn, t1 - kGaeUnroll, R, N);
#pragma unroll
            for (int k = 0; k < kGaeUnroll; ++k) {
                const int t = t1 - 1 - k;
                if (t < 0) break;
                const float r = (float)cur.r[k];
                const float m = cur.d[k] ? 0.0f : 1.0f;  // masks[t+1] = 1 - done[t] (main_player.py:254-258)
                const float dn0 = denorm(cur.v[k], p.vn_std, p.vn_mean);
                float out;
                if (p.use_gae) {
                    // delta = r + gamma * denorm(v[t+1]) * mask - denorm(v[t]);  gae = delta + gamma*lambda * mask * gae
                    const float delta = __fadd_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(p.gamma, dn_next), m)), -dn0);
                    gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(p.gl, m), gae));
                    out = __fadd_rn(gae, dn0);
                } else {
                    // returns[t] = returns[t+1] * gamma * mask + r
                    out = __fadd_rn(__fmul_rn(__fmul_rn(ret_next, p.gamma), m), r);
                    ret_next = out;
                }
                __stcs(ret + (size_t)t * R, out);
                const float a = __fadd_rn(out, -dn0);
                if (adv) __stcs(adv + (size_t)t * R, a);
                s1 += (double)a, s2 += (double)a * (double)a;
                dn_next = dn0;
            }
            cur = nxt;
        }
    }
    if (p.stats != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o);
            s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
        }
        __shared__ double sh[2][kGaeThreads / 32];
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) sh[0][w] = s1, sh[1][w] = s2;
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, b = 0.0;
            for (int i = 0; i < kGaeThreads / 32; ++i) a += sh[0][i], b += sh[1][i];
            atomicAdd(p.stats + 0, a);
            atomicAdd(p.stats + 1, b);
            const int cols = min(kGaeThreads, p.R - (int)blockIdx.x * kGaeThreads);
            atomicAdd(p.stats + 2, (double)cols * (double)p.T);
        }
    }
}

// (a - mean) / (std + 1e-5) with the unbiased std torch.Tensor.std() computes (r_mappo.py:180-182)
__global__ void normalize_adv_kernel(float* adv, size_t n, const double* stats) {
    const double cnt = stats[2];
    const double mean = stats[0] / cnt;
    double var = (stats[1] - cnt * mean * mean) / (cnt - 1.0);
    var = var > 0.0 ? var : 0.0;
    const float meanf = (float)mean, inv = 1.0f / ((float)sqrt(var) + 1e-5f);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4* adv4 = reinterpret_cast<float4*>(adv);
    const size_t n4 = ((reinterpret_cast<uintptr_t>(adv) & 15u) == 0) ? n / 4 : 0;
    for (size_t j = i; j < n4; j += stride) {
        float4 a = adv4[j];
        a.x = (a.x - meanf) * inv, a.y = (a.y - meanf) * inv, a.z = (a.z - meanf) * inv, a.w = (a.w - meanf) * inv;
        adv4[j] = a;
    }
    for (size_t j = n4 * 4 + i; j < n; j += stride) adv[j] = (adv[j] - meanf) * inv;
}

}  // namespace

extern "C" int ocb_compute_returns(int device, const ocb_returns_cfg* cfg, int T, int P, int N, const float* value_preds,
                                   const int32_t* rewards, const int32_t* done, float* returns, float* advantages,
                                   double* adv_stats, void* stream) {
    if (cfg == nullptr || value_preds == nullptr || rewards == nullptr || done == nullptr || returns == nullptr)
        return fail(OCB_ERR_INVALID_ARG, "NULL argument");
    if (cfg->struct_size != sizeof(ocb_returns_cfg)) return fail(OCB_ERR_INVALID_ARG, "ocb_returns_cfg ABI mismatch");
    if (T < 1 || P < 1 || N < 1 || (long long)P * N > 0x7FFFFFFFLL) return fail(OCB_ERR_INVALID_ARG, "bad T / P / N");
    const int ndev = ocb_device_count();
    if (ndev <= 0) return fail(OCB_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(OCB_ERR_INVALID_ARG, "device %d not in 0..%d", device, ndev - 1);
    DeviceGuard guard(device);
    GaeParams p;
    p.value_preds = value_preds, p.rewards = rewards, p.done = done, p.returns = returns, p.advantages = advantages;
    p.stats = adv_stats, p.T = T, p.R = P * N, p.N = N;
    p.gamma = (float)cfg->gamma, p.gl = (float)(cfg->gamma * cfg->gae_lambda);
    p.vn_mean = cfg->vn_mean, p.vn_std = cfg->vn_std, p.use_gae = cfg->use_gae;
    cudaStream_t s = (cudaStream_t)stream;
    if (adv_stats != nullptr) {
        cudaError_t err = cudaMemsetAsync(adv_stats, 0, 3 * sizeof(double), s);
        if (err != cudaSuccess) return fail(OCB_ERR_CUDA, "ocb_compute_returns: %s", cudaGetErrorString(err));
    }
    gae_kernel<<<(p.R + kGaeThreads - 1) / kGaeThreads, kGaeThreads, 0, s>>>(p);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(OCB_ERR_CUDA, "gae kernel launch failed: %s", cudaGetErrorString(err));
    return OCB_OK;
}

extern "C" int ocb_normalize_advantages(int device, float* advantages, size_t n, const double* adv_stats, void* stream) {
    if (advantages == nullptr || adv_stats == nullptr) return fail(OCB_ERR_INVALID_ARG, "NULL argument");
    if (n < 2) return fail(OCB_ERR_INVALID_ARG, "need at least two advantages");
    const int ndev = ocb_device_count();
    if (ndev <= 0) return fail(OCB_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(OCB_ERR_INVALID_ARG, "device %d not in 0..%d", device, ndev - 1);
    DeviceGuard guard(device);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks > (size_t)sms * 8) blocks = (size_t)sms * 8;
    if (blocks < 1) blocks = 1;
    normalize_adv_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(advantages, n, adv_stats);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(OCB_ERR_CUDA, "normalize kernel launch failed: %s", cudaGetErrorString(err));
    return OCB_OK;
}
