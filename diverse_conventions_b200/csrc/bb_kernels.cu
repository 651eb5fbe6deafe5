// bb_kernels.cu — Balance-Beam toy environment (kernels + C ABI).
//
// Semantics: PantheonLine (envs/balance_beam_env.py:95-152, "B:" below); the Madrona
// twin is src/balance_beam_env/sim.cpp:46-155.  One thread owns one world for all K
// steps of a launch; the 7-int observations of the 32 worlds of a warp are transposed
// through shared memory so that every global store is a full 128-byte line.
// Reset positions come from the counter-based RNG keyed by (seed, world, episode):
// the reference's reset uses numpy's global RNG and is not bit-comparable
// (SURVEY.md section 8c), transitions / rewards / dones are.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <new>

#include "api_common.h"
#include "oc_core.cuh"
#include "ocb.h"

using namespace ocb;

namespace {

constexpr int kSpaces = 5, kBuffer = 2, kTime = 3;  // B:13-16
constexpr uint64_t kResetStream = 0xBA1A9CEull;

struct BBWorld {
    int loc[2];
    int time;
    int hist[2][2];  // [player][0]=t-1, [1]=t-2, already +BUFFER shifted (0 = not visited)
    uint32_t episode;
};

// state word: loc0 | loc1<<3 | time<<6 | h0a<<8 | h0b<<12 | h1a<<16 | h1b<<20
__host__ __device__ inline uint32_t bb_pack(const BBWorld& w) {
    return (uint32_t)w.loc[0] | ((uint32_t)w.loc[1] << 3) | ((uint32_t)w.time << 6) | ((uint32_t)w.hist[0][0] << 8) |
           ((uint32_t)w.hist[0][1] << 12) | ((uint32_t)w.hist[1][0] << 16) | ((uint32_t)w.hist[1][1] << 20);
}
__host__ __device__ inline void bb_unpack(uint32_t s, BBWorld& w) {
    w.loc[0] = s & 7, w.loc[1] = (s >> 3) & 7, w.time = (s >> 6) & 3;
    w.hist[0][0] = (s >> 8) & 15, w.hist[0][1] = (s >> 12) & 15;
    w.hist[1][0] = (s >> 16) & 15, w.hist[1][1] = (s >> 20) & 15;
}

// `r` / `r_block`: the reset stream's current Philox block, kept by the caller — one block serves kStepsPerBlock (4)
// consecutive episodes, and episodes last two or three steps, so recomputing it on every reset was a fifth of the kernel's
// instructions and its longest dependent chain
__device__ inline void bb_reset_world(BBWorld& w, uint64_t seed, uint32_t world, ActionRng<2>& r, uint32_t& r_block) {  // B:141-149
    const uint32_t block = w.episode / ActionRng<2>::kStepsPerBlock;
    if (block != r_block) {
        r.refill(seed ^ kResetStream, world, (uint64_t)w.episode);
        r_block = block;
    }
    w.loc[0] = r.action((uint64_t)w.episode, 0, kSpaces);
    w.loc[1] = r.action((uint64_t)w.episode, 1, kSpaces);
    w.time = kTime - 1;
    w.hist[0][0] = w.hist[0][1] = w.hist[1][0] = w.hist[1][1] = 0;
    w.episode += 1;
}

// The reset stream inside the K-step loop.  Lanes end their (one- to three-step) episodes at different steps, so with the
// cached block above some lane of the warp needed a new Philox block on nearly every step and the whole warp walked
// through the ten rounds each time.  Here every lane keeps the block it draws from (`cur`) AND computes the next one a few
// rounds per step (`advance`, the same straight-line code on every lane and step, independent of the transition's chain):
// a block serves four episodes of one or two steps each (seven steps on average), two rounds per step complete the next
// block within five, and the rare lane that gets there earlier (four one-step episodes in a row) runs the missing rounds
// in `reset`.  Same words as ActionRng::refill.
struct BBResetStream {
    static constexpr int kRoundsPerStep = 2;
    uint32_t cur[4], nxt[4];
    uint32_t cur_block, k0, k1;  // block index of `cur` (`nxt` is cur_block + 1); Philox key of nxt's next round
    int rounds;                  // rounds of `nxt` done so far

    __device__ __forceinline__ void start_next(uint64_t seed, uint32_t world) {
        const uint64_t s = seed ^ kResetStream;
        const uint64_t block = (uint64_t)cur_block + 1;
        nxt[0] = world, nxt[1] = (uint32_t)block, nxt[2] = (uint32_t)(block >> 32), nxt[3] = 0u;
        k0 = (uint32_t)s, k1 = (uint32_t)(s >> 32);
        rounds = 0;
    }
    __device__ __forceinline__ void round() {
        const uint32_t hi0 = __umulhi(0xD2511F53u, nxt[0]), lo0 = 0xD2511F53u * nxt[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, nxt[2]), lo1 = 0xCD9E8D57u * nxt[2];
        const uint32_t n0 = hi1 ^ nxt[1] ^ k0, n2 = hi0 ^ nxt[3] ^ k1;
        nxt[0] = n0, nxt[1] = lo1, nxt[2] = n2, nxt[3] = lo0;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
        ++rounds;
    }
    __device__ __forceinline__ void advance() {
#pragma unroll
        for (int q = 0; q < kRoundsPerStep; ++q)
            if (rounds < 10) round();
    }
    __device__ __forceinline__ void init(uint64_t seed, uint32_t world, uint32_t episode) {
        ActionRng<2> r;
        r.refill(seed ^ kResetStream, world, (uint64_t)episode);
        cur_block = episode / ActionRng<2>::kStepsPerBlock;
#pragma unroll
        for (int q = 0; q < 4; ++q) cur[q] = r.r[q];
        start_next(seed, world);
        while (rounds < 10) round();
    }
    // B:141-149 for the lanes with `d` set, written as selects: some lane of a warp ends an episode on nearly every step, so a
    // branch would make the whole warp walk the reset code anyway, plus the divergence bookkeeping
    __device__ __forceinline__ void reset_if(BBWorld& w, bool d, uint64_t seed, uint32_t world) {
        const uint32_t block = w.episode / ActionRng<2>::kStepsPerBlock;
        const bool swap = d && block != cur_block;  // == cur_block + 1: episodes count up by one
        if (swap && rounds < 10) {  // rare (four one-step episodes in a row): the next block is not complete yet
            while (rounds < 10) round();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) cur[q] = swap ? nxt[q] : cur[q];
        cur_block = swap ? block : cur_block;
        {   // start_next for the swapping lanes
            const uint64_t s = seed ^ kResetStream;
            const uint64_t nb = (uint64_t)cur_block + 1;
            nxt[0] = swap ? world : nxt[0], nxt[1] = swap ? (uint32_t)nb : nxt[1], nxt[2] = swap ? (uint32_t)(nb >> 32) : nxt[2];
            nxt[3] = swap ? 0u : nxt[3];
            k0 = swap ? (uint32_t)s : k0, k1 = swap ? (uint32_t)(s >> 32) : k1;
            rounds = swap ? 0 : rounds;
        }
        const int h = (int)(w.episode % ActionRng<2>::kStepsPerBlock) * 2;  // 16-bit slices h, h + 1 of the block
        const uint32_t lo = (h & 2) ? cur[1] : cur[0], hi = (h & 2) ? cur[3] : cur[2];
        const uint32_t word = (h & 4) ? hi : lo;
        w.loc[0] = d ? (int)(((word & 0xFFFFu) * (uint32_t)kSpaces) >> 16) : w.loc[0];
        w.loc[1] = d ? (int)(((word >> 16) * (uint32_t)kSpaces) >> 16) : w.loc[1];
        w.time = d ? kTime - 1 : w.time;
        w.hist[0][0] = d ? 0 : w.hist[0][0], w.hist[0][1] = d ? 0 : w.hist[0][1];
        w.hist[1][0] = d ? 0 : w.hist[1][0], w.hist[1][1] = d ? 0 : w.hist[1][1];
        w.episode += d ? 1u : 0u;
    }
    // the same with a branch (one-off resets outside the step loop)
    __device__ __forceinline__ void reset(BBWorld& w, uint64_t seed, uint32_t world) {
        const uint32_t block = w.episode / ActionRng<2>::kStepsPerBlock;
        if (block != cur_block) {  // == cur_block + 1: episodes count up by one
            while (rounds < 10) round();
#pragma unroll
            for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
            cur_block = block;
            start_next(seed, world);
        }
        const int h = (int)(w.episode % ActionRng<2>::kStepsPerBlock) * 2;  // 16-bit slices h, h + 1 of the block
        const uint32_t lo = (h & 2) ? cur[1] : cur[0], hi = (h & 2) ? cur[3] : cur[2];
        const uint32_t word = (h & 4) ? hi : lo;
        w.loc[0] = (int)(((word & 0xFFFFu) * (uint32_t)kSpaces) >> 16);
        w.loc[1] = (int)(((word >> 16) * (uint32_t)kSpaces) >> 16);
        w.time = kTime - 1;
        w.hist[0][0] = w.hist[0][1] = w.hist[1][0] = w.hist[1][1] = 0;
        w.episode += 1;
    }
};

__device__ inline int bb_move(int a) { return a - 2 + (a >> 1); }  // B:14: [-2, -1, 1, 2][a], a in 0..3, without branches

// Rewards (B:129-137) are Python floats (fp64) narrowed to float32 by the tensor they are written into.  For every value
// the env can produce — |loc0 - loc1| in 0..8 and the out-of-bounds penalty -SPACES * (time + 1) * 0.2 — the float32
// product is the same float, so the kernel multiplies in float32 (the fp64 multiply and its two conversions were 4 % of the
// kernel's instructions); checked here at compile time.
constexpr bool bb_rewards_match_fp64() {
    for (int d = 1; d <= 8; ++d)
        if ((float)(-(double)d * 0.2) != -(float)d * 0.2f) return false;
    for (int t = 0; t <= kTime; ++t)
        if ((float)((double)(-kSpaces * (t + 1)) * 0.2) != (float)(-kSpaces * (t + 1)) * 0.2f) return false;
    return true;
}
static_assert(bb_rewards_match_fp64(), "float32 rewards differ from the reference's fp64-then-narrowed ones");

struct BBParams {
    uint32_t* state;
    uint32_t* episode;
    int N, K;
    unsigned long long seed, step0;
    unsigned int world0;
    const int32_t* actions;  // [K][2][N] or nullptr -> RNG
    uint8_t* actions_out;
    int32_t* obs;  // [K][2][N][7]
    float* rew;    // [K][2][N]
    int32_t* done; // [K][N]
    int mode;      // 0 rollout, 1 observe only, 2 reset then observe
};

__device__ inline void bb_write_obs(const BBWorld& w, int32_t* tile /* [2][32*7] */, int lane) {  // B:116-121
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        int32_t* o = tile + v * 224 + lane * 7;
        const int me = v, ot = 1 - v;
        o[0] = w.loc[me] + kBuffer, o[1] = w.hist[me][0], o[2] = w.hist[me][1];
        o[3] = w.loc[ot] + kBuffer, o[4] = w.hist[ot][0], o[5] = w.hist[ot][1];
        o[6] = w.time;
    }
}

// `vec`: full tile and 16-byte aligned rows (N % 4 == 0): the 224 ints of a view leave as 56 16-byte stores
// `obs`: this tile's rows of view 0 of the step; view 1 follows N rows later
__device__ inline void bb_flush_obs(const int32_t* tile, int32_t* obs, int N, int nvalid, int lane, bool vec) {
    if (vec) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            uint4* dst = reinterpret_cast<uint4*>(obs + (size_t)v * N * 7);
            const uint4* src = reinterpret_cast<const uint4*>(tile + v * 224);
            __stcs(dst + lane, src[lane]);
            if (lane < 24) __stcs(dst + 32 + lane, src[32 + lane]);
        }
        return;
    }
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        int32_t* dst = obs + (size_t)v * N * 7;
        const int cnt = nvalid * 7;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int idx = j * 32 + lane;
            if (idx < cnt) __stcs(dst + idx, tile[v * 224 + idx]);
        }
    }
}

__global__ void __launch_bounds__(256) bb_kernel(const BBParams prm) {
    __shared__ __align__(16) int32_t tiles[8][2 * 224];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n0 = (blockIdx.x * 8 + warp) * 32;
    const int N = prm.N;
    if (n0 >= N) return;
    const int nvalid = min(32, N - n0);
    const int n = n0 + lane;
    const bool valid = n < N;
    const int nl = valid ? n : N - 1;
    int32_t* tile = tiles[warp];

    BBWorld w;
    bb_unpack(prm.state[nl], w);
    w.episode = prm.episode[nl];
    const uint32_t gworld = prm.world0 + (uint32_t)nl;
    ActionRng<2> reset_rng;
    uint32_t reset_block = 0xFFFFFFFFu;  // no block cached (episode counters stay far below 2^34)
    const bool vec = nvalid == 32 && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(prm.obs) & 15u) == 0;

    if (prm.mode != 0) {
        if (prm.mode == 2) bb_reset_world(w, prm.seed, gworld, reset_rng, reset_block);
        if (prm.obs != nullptr) {
            bb_write_obs(w, tile, lane);
            __syncwarp();
            bb_flush_obs(tile, prm.obs + (size_t)n0 * 7, N, nvalid, lane, vec);
        }
    } else {
        ActionRng<2> rng;
        unsigned long long t = prm.step0;
        const bool use_rng = prm.actions == nullptr;
        if (use_rng && (t % ActionRng<2>::kStepsPerBlock) != 0) rng.refill(prm.seed, gworld, t);
        BBResetStream rs;
        rs.init(prm.seed, gworld, w.episode);
        // per-lane output cursors, advanced by one step's stride each iteration (no 64-bit multiplies in the loop)
        const size_t PN = 2 * (size_t)N;
        const int32_t* act_ptr = use_rng ? nullptr : prm.actions + nl;
        uint8_t* aout_ptr = (prm.actions_out != nullptr && valid) ? prm.actions_out + n : nullptr;
        float* rew_ptr = (prm.rew != nullptr && valid) ? prm.rew + n : nullptr;
        int32_t* done_ptr = (prm.done != nullptr && valid) ? prm.done + n : nullptr;
        int32_t* obs_ptr = prm.obs != nullptr ? prm.obs + (size_t)n0 * 7 : nullptr;
        for (int k = 0; k < prm.K; ++k, ++t) {
            rs.advance();
            int a0, a1;
            if (use_rng) {
                if ((t % ActionRng<2>::kStepsPerBlock) == 0) rng.refill(prm.seed, gworld, t);
                a0 = rng.action(t, 0, 4), a1 = rng.action(t, 1, 4);
            } else {
                a0 = act_ptr[0] & 3, a1 = act_ptr[N] & 3;
                act_ptr += PN;
            }
            if (aout_ptr != nullptr) {
                aout_ptr[0] = (uint8_t)a0, aout_ptr[N] = (uint8_t)a1;
                aout_ptr += PN;
            }
            // B:123-139
            w.hist[0][1] = w.hist[0][0], w.hist[0][0] = w.loc[0] + kBuffer;
            w.hist[1][1] = w.hist[1][0], w.hist[1][0] = w.loc[1] + kBuffer;
            w.loc[0] += bb_move(a0);
            w.loc[1] += bb_move(a1);
            w.time -= 1;
            const int diff = abs(w.loc[0] - w.loc[1]);
            // (unsigned compare: loc < 0 wraps above SPACES)
            const bool oob = (unsigned)w.loc[0] >= (unsigned)kSpaces || (unsigned)w.loc[1] >= (unsigned)kSpaces;
            const bool d = (w.time == 0) || oob;
            float r = (diff == 0) ? 1.0f : -(float)diff * 0.2f;  // == the reference's fp64 value narrowed (static_assert above)
            r = oob ? (float)(-kSpaces * (w.time + 1)) * 0.2f : r;
            rs.reset_if(w, d, prm.seed, gworld);  // pantheonrl_extension/vectorenv.py:369-370
            if (rew_ptr != nullptr) {
                rew_ptr[0] = r, rew_ptr[N] = r;
                rew_ptr += PN;
            }
            if (done_ptr != nullptr) {
                *done_ptr = d ? 1 : 0;
                done_ptr += N;
            }
            if (obs_ptr != nullptr) {
                bb_write_obs(w, tile, lane);
                __syncwarp();
                bb_flush_obs(tile, obs_ptr, N, nvalid, lane, vec);
                __syncwarp();
                obs_ptr += PN * 7;
            }
        }
    }
    if (valid && prm.mode != 1) {
        prm.state[n] = bb_pack(w);
        prm.episode[n] = w.episode;
    }
}

__global__ void bb_export_kernel(const uint32_t* state, const uint32_t* episode, int32_t* out, int N) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        BBWorld w;
        bb_unpack(state[n], w);
        int32_t* o = out + (size_t)n * 8;
        o[0] = w.loc[0], o[1] = w.loc[1], o[2] = w.time;
        o[3] = w.hist[0][0], o[4] = w.hist[0][1], o[5] = w.hist[1][0], o[6] = w.hist[1][1];
        o[7] = (int32_t)episode[n];
    }
}

__global__ void bb_import_kernel(const int32_t* in, uint32_t* state, uint32_t* episode, int N, int* bad) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const int32_t* o = in + (size_t)n * 8;
        bool ok = o[0] >= 0 && o[0] < kSpaces && o[1] >= 0 && o[1] < kSpaces && o[2] >= 1 && o[2] <= kTime - 1;
        for (int j = 3; j < 7; ++j) ok = ok && o[j] >= 0 && o[j] < 16;
        if (!ok) {
            atomicAdd(bad, 1);
            continue;
        }
        BBWorld w;
        w.loc[0] = o[0], w.loc[1] = o[1], w.time = o[2];
        w.hist[0][0] = o[3], w.hist[0][1] = o[4], w.hist[1][0] = o[5], w.hist[1][1] = o[6];
        state[n] = bb_pack(w);
        episode[n] = (uint32_t)o[7];
    }
}

}  // namespace

struct bb_env {
    int device;
    int N;
    uint64_t seed, step_count;
    uint32_t world0;
    uint32_t* d_state;
    uint32_t* d_episode;
};

static int bb_launch(bb_env* e, BBParams& p, void* stream) {
    p.state = e->d_state, p.episode = e->d_episode, p.N = e->N;
    p.seed = e->seed, p.step0 = e->step_count, p.world0 = e->world0;
    const int ctas = (e->N + 255) / 256;
    bb_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(p);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(OCB_ERR_CUDA, "bb kernel launch failed: %s", cudaGetErrorString(err));
    return OCB_OK;
}

extern "C" int bb_destroy(bb_env* e) {
    if (e == nullptr) return OCB_OK;
    DeviceGuard guard(e->device);
    CaptureRelaxed relaxed;  // safe while another stream is being captured
    cudaFree(e->d_state);
    cudaFree(e->d_episode);
    delete e;
    return OCB_OK;
}

extern "C" int bb_create(int device, uint32_t num_worlds, uint64_t seed, bb_env** out) {
    if (out == nullptr) return fail(OCB_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (num_worlds < 1 || num_worlds > (1u << 30)) return fail(OCB_ERR_INVALID_ARG, "num_worlds out of range");
    const int ndev = ocb_device_count();
    if (ndev <= 0) return fail(OCB_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(OCB_ERR_INVALID_ARG, "device %d not in 0..%d", device, ndev - 1);
    bb_env* e = new (std::nothrow) bb_env();
    if (e == nullptr) return fail(OCB_ERR_INVALID_ARG, "out of host memory");
    memset(e, 0, sizeof(*e));
    e->device = device, e->N = (int)num_worlds, e->seed = seed;
    DeviceGuard guard(device);
    cudaError_t err = cudaMalloc(&e->d_state, sizeof(uint32_t) * (size_t)num_worlds);
    if (err == cudaSuccess) err = cudaMalloc(&e->d_episode, sizeof(uint32_t) * (size_t)num_worlds);
    if (err == cudaSuccess) err = cudaMemset(e->d_state, 0, sizeof(uint32_t) * (size_t)num_worlds);
    if (err == cudaSuccess) err = cudaMemset(e->d_episode, 0, sizeof(uint32_t) * (size_t)num_worlds);
    if (err != cudaSuccess) {
        cudaGetLastError();
        bb_destroy(e);
        return fail(OCB_ERR_CUDA, "bb_create: %s", cudaGetErrorString(err));
    }
    BBParams p;
    memset(&p, 0, sizeof(p));
    p.mode = 2;
    int rc = bb_launch(e, p, nullptr);
    if (rc == OCB_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(OCB_ERR_CUDA, "bb_create: reset kernel failed");
    if (rc != OCB_OK) {
        bb_destroy(e);
        return rc;
    }
    *out = e;
    return OCB_OK;
}

extern "C" int bb_num_worlds(const bb_env* e) { return e ? e->N : fail(OCB_ERR_INVALID_ARG, "env is NULL"); }

extern "C" int bb_reset(bb_env* e, int32_t* obs, void* stream) {
    if (e == nullptr) return fail(OCB_ERR_INVALID_ARG, "env is NULL");
    DeviceGuard guard(e->device);
    BBParams p;
    memset(&p, 0, sizeof(p));
    p.mode = 2, p.obs = obs;
    return bb_launch(e, p, stream);
}

extern "C" int bb_observe(bb_env* e, int32_t* obs, void* stream) {
    if (e == nullptr || obs == nullptr) return fail(OCB_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard guard(e->device);
    BBParams p;
    memset(&p, 0, sizeof(p));
    p.mode = 1, p.obs = obs;
    return bb_launch(e, p, stream);
}

static int bb_run(bb_env* e, int K, const int32_t* actions, int32_t* obs, float* rew, int32_t* done, uint8_t* actions_out,
                  void* stream) {
    if (e == nullptr) return fail(OCB_ERR_INVALID_ARG, "env is NULL");
    if (K < 0) return fail(OCB_ERR_INVALID_ARG, "K must be >= 0");
    DeviceGuard guard(e->device);
    BBParams p;
    memset(&p, 0, sizeof(p));
    p.mode = 0, p.K = K, p.actions = actions, p.actions_out = actions_out, p.obs = obs, p.rew = rew, p.done = done;
    int rc = bb_launch(e, p, stream);
    if (rc == OCB_OK) e->step_count += (uint64_t)K;
    return rc;
}

extern "C" int bb_step(bb_env* e, const int32_t* actions, int32_t* obs, float* reward, int32_t* done, void* stream) {
    if (actions == nullptr) return fail(OCB_ERR_INVALID_ARG, "actions is NULL");
    return bb_run(e, 1, actions, obs, reward, done, nullptr, stream);
}

extern "C" int bb_rollout_random(bb_env* e, int K, int32_t* obs_slab, float* reward, int32_t* done, uint8_t* actions_out,
                                 void* stream) {
    return bb_run(e, K, nullptr, obs_slab, reward, done, actions_out, stream);
}

extern "C" int bb_get_state(bb_env* e, int32_t* h_state, size_t n_ints) {
    if (e == nullptr || h_state == nullptr) return fail(OCB_ERR_INVALID_ARG, "NULL argument");
    const size_t need = (size_t)e->N * 8;
    if (n_ints != need) return fail(OCB_ERR_INVALID_ARG, "state buffer has %zu ints, expected %zu", n_ints, need);
    DeviceGuard guard(e->device);
    int32_t* d = nullptr;
    cudaError_t err = cudaDeviceSynchronize();
    if (err == cudaSuccess) err = cudaMalloc(&d, need * sizeof(int32_t));
    if (err == cudaSuccess) {
        bb_export_kernel<<<(e->N + 255) / 256, 256>>>(e->d_state, e->d_episode, d, e->N);
        err = cudaMemcpy(h_state, d, need * sizeof(int32_t), cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return fail(OCB_ERR_CUDA, "bb_get_state: %s", cudaGetErrorString(err));
    }
    return OCB_OK;
}

extern "C" int bb_set_state(bb_env* e, const int32_t* h_state, size_t n_ints) {
    if (e == nullptr || h_state == nullptr) return fail(OCB_ERR_INVALID_ARG, "NULL argument");
    const size_t need = (size_t)e->N * 8;
    if (n_ints != need) return fail(OCB_ERR_INVALID_ARG, "state buffer has %zu ints, expected %zu", n_ints, need);
    DeviceGuard guard(e->device);
    int32_t* d = nullptr;
    int* d_bad = nullptr;
    uint32_t *t_state = nullptr, *t_episode = nullptr;
    int bad = 0;
    const size_t N = e->N;
    cudaError_t err = cudaDeviceSynchronize();
    if (err == cudaSuccess) err = cudaMalloc(&d, need * sizeof(int32_t));
    if (err == cudaSuccess) err = cudaMalloc(&d_bad, sizeof(int));
    if (err == cudaSuccess) err = cudaMalloc(&t_state, sizeof(uint32_t) * N);
    if (err == cudaSuccess) err = cudaMalloc(&t_episode, sizeof(uint32_t) * N);
    if (err == cudaSuccess) err = cudaMemset(d_bad, 0, sizeof(int));
    if (err == cudaSuccess) err = cudaMemcpy(d, h_state, need * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (err == cudaSuccess) {
        bb_import_kernel<<<(e->N + 255) / 256, 256>>>(d, t_state, t_episode, e->N, d_bad);
        err = cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
    }
    if (err == cudaSuccess && bad == 0) {
        err = cudaMemcpy(e->d_state, t_state, sizeof(uint32_t) * N, cudaMemcpyDeviceToDevice);
        if (err == cudaSuccess) err = cudaMemcpy(e->d_episode, t_episode, sizeof(uint32_t) * N, cudaMemcpyDeviceToDevice);
    }
    cudaFree(d), cudaFree(d_bad), cudaFree(t_state), cudaFree(t_episode);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return fail(OCB_ERR_CUDA, "bb_set_state: %s", cudaGetErrorString(err));
    }
    if (bad) return fail(OCB_ERR_BAD_STATE, "%d world(s) hold an invalid Balance-Beam state", bad);
    return OCB_OK;
}
