// oc_kernels.cu — sm_100a kernels of the batched Overcooked simulator.
//
// oc_rollout_kernel<P,G>: K fused environment steps per launch.
//   * a warp owns a tile of WPW = 32/G consecutive worlds for the whole launch;
//     every world is served by G lanes that execute the (sequential, branchy)
//     transition redundantly — each lane on its own registers and its own private
//     column of the cell objects in shared memory, so the lanes of a world never
//     read-modify-write a shared word (no reliance on warp lock-step; racecheck
//     clean) — and split the observation byte pokes (disjoint bytes);
//   * world state lives in registers (players) and shared memory (cell objects)
//     for all K steps; HBM is touched only for the mandatory I/O:
//     actions in, observation planes / reward / done out;
//   * each tile keeps its P x WPW observation planes resident in shared memory,
//     laid out exactly like the [P, N, W, H, C] output so that one view of the
//     tile is one contiguous run of WPW*S*C bytes in HBM.  After the few bytes
//     touched by a transition are rewritten the run is streamed out with one TMA
//     bulk copy (cp.async.bulk.global.shared::cta) per view, or with 16-byte
//     coalesced streaming stores when the run is not 16-byte aligned.
// See DESIGN.md for the byte accounting and the roofline.
#include <cuda_runtime.h>
#include <stdint.h>

#include "oc_core.cuh"
#include "oc_device.cuh"
#include "oc_kernels.h"

namespace ocb {

// shared-memory carve-up of one CTA (all offsets 16-byte aligned):
//   Tables | template[SC] | mbarriers | per warp: planes[P][view_stride] , objs[S][32] u16 (one column per LANE)
constexpr int kFillBarBytes = 64;
template <int P, int G>
struct Carve {
    static constexpr int WPW = 32 / G;
    int view_stride;
    size_t warp_bytes, warp0, bars;
    __device__ Carve(int S, int SC) {
        view_stride = (int)align16((size_t)WPW * SC);
        warp_bytes = (size_t)P * view_stride + align16((size_t)S * 32 * 2);
        bars = align16(sizeof(Tables)) + align16((size_t)SC);  // one mbarrier per warp (tile_fill_begin)
        warp0 = bars + kFillBarBytes;
    }
};

__device__ __forceinline__ void stage_tables(uint8_t* smem, const RolloutParams& prm) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(prm.tables);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    const int SC = reinterpret_cast<const Tables*>(smem)->SC;
    uint8_t* tmpl = smem + align16(sizeof(Tables));
    if ((SC & 3) == 0) {  // the device template is 16-byte aligned
        const uint32_t* t4 = reinterpret_cast<const uint32_t*>(prm.tmpl);
        for (int i = threadIdx.x; i < (SC >> 2); i += blockDim.x) reinterpret_cast<uint32_t*>(tmpl)[i] = t4[i];
    } else {
        for (int i = threadIdx.x; i < SC; i += blockDim.x) tmpl[i] = prm.tmpl[i];
    }
    __syncthreads();
}

// kFill: rebuild the planes of the tile at the start of the launch by one bulk copy per view (single-step launches, where
// that rebuild is most of the work; the K-step launches keep the per-lane copy of their first step: their step loop runs
// at the register limit of its launch bounds and measured 1-2 % slower with the extra state around it)
template <int P, int G, bool kFill>
__global__ void __launch_bounds__(kThreadsPerCta, (G == 8 ? 6 : 4)) oc_rollout_kernel(const RolloutParams prm) {
    constexpr int WPW = 32 / G;
    extern __shared__ __align__(16) uint8_t smem[];
    stage_tables(smem, prm);
    const Tables& tb = *reinterpret_cast<const Tables*>(smem);
    const Consts c = load_consts(tb);
    const int SC = tb.SC, S = tb.S;
    const uint8_t* tmpl = smem + align16(sizeof(Tables));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wi = lane / G, g = lane % G;
    const Carve<P, G> cv(S, SC);
    const int view_stride = cv.view_stride;
    uint8_t* planes = smem + cv.warp0 + warp * cv.warp_bytes;                                 // [P][WPW][SC]
    uint16_t* myobjs = reinterpret_cast<uint16_t*>(planes + (size_t)P * view_stride) + lane;  // [S][32], private column
    uint8_t* myplanes = planes + wi * SC;                                                     // + v*view_stride

    const int N = prm.N;
    const int n0 = (blockIdx.x * (blockDim.x >> 5) + warp) * WPW;
    if (n0 >= N) return;  // whole warp idle (no block-level sync below)
    const int nvalid = min(WPW, N - n0);
    const int n = n0 + wi;
    const bool valid = n < N;
    const int nl = valid ? n : N - 1;

    // planes <- static template of the whole tile, by bulk copy, while the state loads and the first transition runs
    const bool want_obs0 = prm.obs != nullptr;
    const int tile_bytes = WPW * SC;
    const bool tile_fill = kFill && want_obs0 && (tile_bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(prm.tmpl) & 15u) == 0;
    const uint32_t fill_bar = smem_u32(smem + cv.bars) + 8u * (uint32_t)warp;
    if (tile_fill && lane == 0) tile_fill_begin(fill_bar, smem_u32(planes), view_stride, prm.tmpl, (uint32_t)tile_bytes, P);
    __syncwarp();

    World<P> w;
    load_world<P, G, kFill ? 8 : 1>(tb, c, prm, nl, g, myobjs, w);
    // (waited for here, not inside the step loop: the loop runs at the register limit of its launch bounds)
    if (tile_fill) tile_fill_wait(fill_bar);
    bool rebuild1 = !tile_fill;  // first step of the launch: phase 1 copies the template unless the bulk copy did
    int cur_return = prm.cur_return[nl];
    long long ret_add = 0;
    int ep_add = 0;

    const bool use_rng = prm.actions == nullptr;
    const bool want_obs = prm.obs != nullptr;
    ActionRng<P> rng;
    unsigned long long t = prm.step0_dev != nullptr ? *prm.step0_dev : prm.step0;
    const uint32_t gworld = prm.world0 + (uint32_t)nl;
    if (use_rng && (t % ActionRng<P>::kStepsPerBlock) != 0) rng.refill(prm.seed, gworld, t);

    // per-lane output cursors, advanced by one step's stride each iteration
    const size_t PN = (size_t)P * N;
    int32_t* rew_ptr = prm.rew ? prm.rew + n : nullptr;      // + i*N
    int32_t* done_ptr = prm.done ? prm.done + n : nullptr;
    uint8_t* aout_ptr = prm.actions_out ? prm.actions_out + n : nullptr;
    size_t act_idx = nl;                                     // + i*N
    int8_t* obs_ptr = want_obs ? prm.obs + (size_t)n0 * SC : nullptr;
    const size_t obs_view_stride = (size_t)N * SC, obs_step_stride = PN * SC;
    const int nbytes = nvalid * SC;
    const bool tma_ok = want_obs && prm.use_tma && ((nbytes & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(obs_ptr) & 15u) == 0) && ((obs_view_stride & 15u) == 0);
    const uint32_t planes_s = smem_u32(planes);
    bool tma_pending = false;

    for (int k = 0; k < prm.K; ++k, ++t) {
        // ---- joint action
        int act[P];
        if (use_rng) {
            if ((t % ActionRng<P>::kStepsPerBlock) == 0) rng.refill(prm.seed, gworld, t);
#pragma unroll
            for (int i = 0; i < P; ++i) act[i] = rng.action(t, i, 6);
        } else {
#pragma unroll
            for (int i = 0; i < P; ++i) act[i] = load_action(prm.actions, prm.act_dtype, act_idx + (size_t)i * N);
            act_idx += PN;
        }
        if (aout_ptr != nullptr) {
            if (valid) {
#pragma unroll
                for (int i = 0; i < P; ++i)
                    if (i % G == g) aout_ptr[(size_t)i * N] = (uint8_t)act[i];
            }
            aout_ptr += PN;
        }

        // ---- transition (registers + shared memory only)
        int oldslot[P];
        uint32_t dirty[P];
#pragma unroll
        for (int i = 0; i < P; ++i) oldslot[i] = w.slot[i];
        uint32_t ticked;
        const int r = step_world<P>(tb, c, w, myobjs, 32, act, dirty, ticked);
        const bool done = w.timestep >= c.horizon;  // envs/overcooked2_env.py:334
        cur_return += r;
        if (done) {  // auto-reset, pantheonrl_extension/vectorenv.py:369-370
            ret_add += cur_return;
            ep_add += 1;
            cur_return = 0;
            reset_world<P>(tb, w);
            for (int idx = 0; idx < c.n_objcells; ++idx) myobjs[(int)tb.objcells[idx] * 32] = 0;  // own column
        }
        if (rew_ptr != nullptr) {
            if (valid) {
#pragma unroll
                for (int i = 0; i < P; ++i)
                    if (i % G == g) rew_ptr[(size_t)i * N] = r;
            }
            rew_ptr += PN;
        }
        if (done_ptr != nullptr) {
            if (valid && g == G - 1) *done_ptr = done ? 1 : 0;
            done_ptr += N;
        }

        // ---- observation planes: rewrite touched bytes, stream the tile out
        if (want_obs) {
            const bool full = (k == 0) || done;
            if (tma_pending) {  // the previous bulk copy must have finished reading the planes
                if (lane == 0) bulk_wait_read_all();
                tma_pending = false;
            }
            __syncwarp();
            obs_phase1<P, G>(tb, myplanes, view_stride, tmpl, rebuild1 || done, g, oldslot);
            rebuild1 = false;
            __syncwarp();
            obs_phase2<P, G>(tb, c, myplanes, view_stride, myobjs, 32, full, g, w, dirty, ticked);
            if (tma_ok) {
                fence_proxy_async_smem();  // generic-proxy pokes -> visible to the async proxy
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int v = 0; v < P; ++v)
                        bulk_store_s2g(obs_ptr + v * obs_view_stride, planes_s + v * view_stride, (uint32_t)nbytes);
                    bulk_commit();
                }
                tma_pending = true;
            } else {
                __syncwarp();
#pragma unroll
                for (int v = 0; v < P; ++v) warp_copy_out(obs_ptr + v * obs_view_stride, planes + v * view_stride, nbytes, lane);
            }
            obs_ptr += obs_step_stride;
        } else {
            __syncwarp();
        }
    }
    if (tma_pending && lane == 0) bulk_wait_read_all();
    // device-side step counter: read by kernels launched after this one (policy sampling offset)
    if (prm.step_counter != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *prm.step_counter += (unsigned long long)prm.K;

    // ---- store world state back
    if (valid) {
#pragma unroll
        for (int i = 0; i < P; ++i)
            if (i % G == g) prm.players[(size_t)i * N + n] = player_pack(w.pos[i], w.orient[i], w.held[i]);
        store_world_objs<G>(tb, c, prm, n, g, myobjs);
        if (g == 0) {
            prm.timestep[n] = w.timestep;
            prm.cur_return[n] = cur_return;
            if (ep_add) {
                prm.ret_sum[n] += ret_add;
                prm.episodes[n] += ep_add;
            }
        }
    }
}

// observation of the current state (no step): full rebuild + stream out
template <int P, int G>
__global__ void __launch_bounds__(kThreadsPerCta) oc_observe_kernel(const RolloutParams prm) {
    constexpr int WPW = 32 / G;
    extern __shared__ __align__(16) uint8_t smem[];
    stage_tables(smem, prm);
    const Tables& tb = *reinterpret_cast<const Tables*>(smem);
    const Consts c = load_consts(tb);
    const int SC = tb.SC, S = tb.S;
    const uint8_t* tmpl = smem + align16(sizeof(Tables));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wi = lane / G, g = lane % G;
    const Carve<P, G> cv(S, SC);
    const int view_stride = cv.view_stride;
    uint8_t* planes = smem + cv.warp0 + warp * cv.warp_bytes;
    uint16_t* myobjs = reinterpret_cast<uint16_t*>(planes + (size_t)P * view_stride) + lane;
    uint8_t* myplanes = planes + wi * SC;

    const int N = prm.N;
    const int n0 = (blockIdx.x * (blockDim.x >> 5) + warp) * WPW;
    if (n0 >= N) return;
    const int nvalid = min(WPW, N - n0);
    const int nl = min(n0 + wi, N - 1);

    const int tile_bytes = WPW * SC;
    const bool tile_fill = (tile_bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(prm.tmpl) & 15u) == 0;
    const uint32_t fill_bar = smem_u32(smem + cv.bars) + 8u * (uint32_t)warp;
    if (tile_fill && lane == 0) tile_fill_begin(fill_bar, smem_u32(planes), view_stride, prm.tmpl, (uint32_t)tile_bytes, P);
    __syncwarp();

    World<P> w;
    load_world<P, G, 8>(tb, c, prm, nl, g, myobjs, w);
    int noslot[P];
    uint32_t nodirty[P];
#pragma unroll
    for (int i = 0; i < P; ++i) noslot[i] = 0, nodirty[i] = 0xFFFFFFFFu;
    if (tile_fill)
        tile_fill_wait(fill_bar);
    else
        obs_phase1<P, G>(tb, myplanes, view_stride, tmpl, true, g, noslot);
    __syncwarp();
    obs_phase2<P, G>(tb, c, myplanes, view_stride, myobjs, 32, true, g, w, nodirty, 0u);
    __syncwarp();
    const int nbytes = nvalid * SC;
#pragma unroll
    for (int v = 0; v < P; ++v)
        warp_copy_out(prm.obs + ((size_t)v * N + n0) * SC, planes + v * view_stride, nbytes, lane);
}

__global__ void oc_reset_kernel(const Tables* __restrict__ tables, uint32_t* players, uint16_t* objs, int32_t* timestep,
                                int32_t* cur_return, int N) {
    const int P = tables->P, S = tables->S;
    const size_t total = (size_t)N * (size_t)(S > P ? S : P);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(idx % (size_t)N), row = (int)(idx / (size_t)N);
        if (row < P) players[(size_t)row * N + n] = player_pack(tables->start_pos[row], 0, 0u);
        if (row < S) objs[(size_t)row * N + n] = 0;
        if (row == 0) {
            timestep[n] = 0;
            cur_return[n] = 0;
        }
    }
}

// packed ABI state (int32 [N, L], include/ocb.h) <-> device structure-of-arrays
__global__ void oc_export_state_kernel(const Tables* __restrict__ tables, const uint32_t* players, const uint16_t* objs,
                                       const int32_t* timestep, int32_t* out, int N) {
    const int P = tables->P, S = tables->S, L = 1 + 6 * P + 4 * S;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        int32_t* row = out + (size_t)n * L;
        row[0] = timestep[n];
        for (int i = 0; i < P; ++i) {
            const uint32_t pw = players[(size_t)i * N + n];
            const uint32_t h = pw >> 16;
            int32_t* pl = row + 1 + 6 * i;
            pl[0] = (int)(pw & 0xFFFu);
            pl[1] = (int)((pw >> 12) & 3u);
            pl[2] = obj_name(h);
            pl[3] = obj_onions(h);
            pl[4] = obj_tomatoes(h);
            pl[5] = h ? obj_tickp1(h) - 1 : 0;
        }
        for (int cell = 0; cell < S; ++cell) {
            const uint32_t o = objs[(size_t)cell * N + n];
            int32_t* oc = row + 1 + 6 * P + 4 * cell;
            oc[0] = obj_name(o);
            oc[1] = obj_onions(o);
            oc[2] = obj_tomatoes(o);
            oc[3] = o ? obj_tickp1(o) - 1 : 0;
        }
    }
}

// counts (through *bad) the worlds holding a state the CUDA path does not represent:
// out-of-range fields, objects on cells that cannot hold one, a pot holding anything but
// a soup, a cooking soup outside a pot, a player off the AIR cells, two players on one cell.
__global__ void oc_import_state_kernel(const Tables* __restrict__ tables, const int32_t* in, uint32_t* players,
                                       uint16_t* objs, int32_t* timestep, int32_t* cur_return, int N, int* bad) {
    const int P = tables->P, S = tables->S, L = 1 + 6 * P + 4 * S;
    Consts c = load_consts(*tables);
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const int32_t* row = in + (size_t)n * L;
        bool ok = row[0] >= 0;
        auto check_obj = [&](const int32_t* o) {
            if (o[0] == O_NONE) return o[1] == 0 && o[2] == 0 && o[3] == 0;
            bool good = o[0] >= O_TOMATO && o[0] <= O_SOUP && o[1] >= 0 && o[2] >= 0 && o[1] + o[2] <= 3 && o[3] >= -1 &&
                        o[3] <= OCB_MAX_COOK_TIME + 1;
            if (good && o[0] != O_SOUP) good = (o[1] == 0 && o[2] == 0 && o[3] == -1);
            return good;
        };
        for (int i = 0; i < P; ++i) {
            const int32_t* pl = row + 1 + 6 * i;
            ok = ok && pl[0] >= 0 && pl[0] < S && pl[1] >= 0 && pl[1] <= 3 && check_obj(pl + 2);
            if (ok) ok = info_terrain(tables->cell_info[pl[0]]) == T_AIR;
            if (ok) {
                const uint32_t h = pl[2] ? obj_make(pl[2], pl[3], pl[4], pl[5]) : 0u;
                players[(size_t)i * N + n] = player_pack(pl[0], pl[1], h);
            }
            for (int j = 0; ok && j < i; ++j) ok = row[1 + 6 * j] != pl[0];  // unreachable in the MDP (collision rule)
        }
        for (int cell = 0; cell < S; ++cell) {
            const int32_t* oc = row + 1 + 6 * P + 4 * cell;
            ok = ok && check_obj(oc);
            uint32_t o = 0u;
            if (ok && oc[0] != O_NONE) {
                const int t = info_terrain(tables->cell_info[cell]);
                o = obj_make(oc[0], oc[1], oc[2], oc[3]);
                ok = (t == T_POT && oc[0] == O_SOUP) ||
                     (t == T_COUNTER && !(oc[0] == O_SOUP && soup_cooking(*tables, c, o)));
            }
            objs[(size_t)cell * N + n] = (uint16_t)o;
        }
        timestep[n] = row[0];
        cur_return[n] = 0;
        if (!ok) atomicAdd(bad, 1);
    }
}

// device step counter += k as a launch of its own: used after launches that READ the counter at their start
// (graph-captured rollouts), where an in-kernel increment by one CTA would race with late-starting CTAs
__global__ void oc_counter_add_kernel(unsigned long long* ctr, unsigned long long k) { *ctr += k; }

cudaError_t launch_counter_add(unsigned long long* counter, unsigned long long k, cudaStream_t stream) {
    oc_counter_add_kernel<<<1, 1, 0, stream>>>(counter, k);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ launchers
template <int P, int G>
static cudaError_t launch_pg(const RolloutParams& prm, int warps_per_cta, size_t smem_bytes, bool observe_only,
                             cudaStream_t stream) {
    constexpr int WPW = 32 / G;
    const int tiles = (prm.N + WPW - 1) / WPW;
    const int ctas = (tiles + warps_per_cta - 1) / warps_per_cta;
    auto kern = observe_only ? oc_observe_kernel<P, G> : (prm.K <= 2 ? oc_rollout_kernel<P, G, true> : oc_rollout_kernel<P, G, false>);
    if (smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
    }
    kern<<<ctas, warps_per_cta * 32, smem_bytes, stream>>>(prm);
    return cudaGetLastError();
}

size_t rollout_smem_bytes(int P, int S, int C, int G, int warps_per_cta) {
    const int WPW = 32 / G;
    const size_t SC = (size_t)S * C;
    const size_t per_warp = (size_t)P * align16(WPW * SC) + align16((size_t)S * 32 * 2);
    return align16(sizeof(Tables)) + align16(SC) + kFillBarBytes + warps_per_cta * per_warp;
}

template <int P>
static cudaError_t launch_p(const RolloutParams& prm, int G, int warps_per_cta, size_t smem_bytes, bool observe_only,
                            cudaStream_t stream) {
    switch (G) {
        case 1: return launch_pg<P, 1>(prm, warps_per_cta, smem_bytes, observe_only, stream);
        case 2: return launch_pg<P, 2>(prm, warps_per_cta, smem_bytes, observe_only, stream);
        case 4: return launch_pg<P, 4>(prm, warps_per_cta, smem_bytes, observe_only, stream);
        default: return launch_pg<P, 8>(prm, warps_per_cta, smem_bytes, observe_only, stream);
    }
}

cudaError_t launch_rollout(const RolloutParams& prm, int P, int G, int warps_per_cta, size_t smem_bytes,
                           bool observe_only, cudaStream_t stream) {
    switch (P) {
        case 1: return launch_p<1>(prm, G, warps_per_cta, smem_bytes, observe_only, stream);
        case 2: return launch_p<2>(prm, G, warps_per_cta, smem_bytes, observe_only, stream);
        case 3: return launch_p<3>(prm, G, warps_per_cta, smem_bytes, observe_only, stream);
        case 4: return launch_p<4>(prm, G, warps_per_cta, smem_bytes, observe_only, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_reset(const Tables* tables, uint32_t* players, uint16_t* objs, int32_t* timestep, int32_t* cur_return,
                         int N, int rows, cudaStream_t stream) {
    const size_t total = (size_t)N * rows;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    oc_reset_kernel<<<blocks, 256, 0, stream>>>(tables, players, objs, timestep, cur_return, N);
    return cudaGetLastError();
}

cudaError_t launch_export_state(const Tables* tables, const uint32_t* players, const uint16_t* objs,
                                const int32_t* timestep, int32_t* out, int N, cudaStream_t stream) {
    oc_export_state_kernel<<<(N + 127) / 128, 128, 0, stream>>>(tables, players, objs, timestep, out, N);
    return cudaGetLastError();
}

cudaError_t launch_import_state(const Tables* tables, const int32_t* in, uint32_t* players, uint16_t* objs,
                                int32_t* timestep, int32_t* cur_return, int N, int* bad, cudaStream_t stream) {
    oc_import_state_kernel<<<(N + 127) / 128, 128, 0, stream>>>(tables, in, players, objs, timestep, cur_return, N, bad);
    return cudaGetLastError();
}

}  // namespace ocb
