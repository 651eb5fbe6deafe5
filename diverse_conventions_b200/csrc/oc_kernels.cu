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
    const int TWD = (prm.tile_worlds > 0 && prm.tile_worlds < WPW) ? prm.tile_worlds : WPW;  // worlds of this warp's tile
    const int n0 = (blockIdx.x * (blockDim.x >> 5) + warp) * TWD;
    if (n0 >= N) return;  // whole warp idle (no block-level sync below)
    const int nvalid = min(TWD, N - n0);
    const int n = n0 + wi;
    const bool valid = wi < nvalid;
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

// ------------------------------------------------------------------ role-split K-step kernel
// oc_rollout_split_kernel<GE, TW>: the same K env steps with the two halves of a step on DIFFERENT warps.
// In oc_rollout_kernel a warp's step is one dependent chain (transition -> plane pokes -> fence -> bulk store, ~4,000
// cycles at one world per lane), and the launch is bound by that latency whenever an SM holds too few worlds to overlap
// enough chains (8,192 worlds: 0.66 of the HBM peak; 16,384 worlds on a box that runs under a power cap: 0.92).
// Here a group of 1 + GE warps owns a tile of 32 worlds:
//   * the TRANSITION warp (one world per lane) keeps the state in registers / its private object columns, draws or reads
//     the actions, steps, writes reward / done / actions and publishes per world-step a 20-byte RECORD in a shared-memory
//     ring: both packed player words (+ the episode-end bit), the two interact targets with their new objects, the
//     objects of the pots whose soup ticked — everything the observation of the new state differs by;
//   * GE ENCODER warps (32/GE worlds each, GE lanes per world, no redundant transition) keep the planes of their worlds
//     resident in shared memory, apply the record (clear the cells the players left, re-encode the touched cells, poke the
//     players) and stream their part of the tile out with one bulk store per view.
// The ring (kSplitDepth steps, full / empty named barriers: a waiting warp is descheduled and takes no issue slots from
// the working ones — with mbarrier poll loops a third of the issued instructions were polls and an SM holding four groups
// ran each of them at half speed) lets the transition run ahead, so a step costs
// max(transition, encode) instead of their sum, and the encode chain is split GE ways.  Encoders build the planes of the
// launch's initial state from the state arrays in HBM on their own (nothing but the ring is shared with the transition
// warp).  Restricted to P = 2 and at most two pots (the record layout); everything else runs oc_rollout_kernel.
constexpr int kSplitDepth = 4;  // ring slots (steps the transition may run ahead)
constexpr int kSplitWords = 5;  // 32-bit words per world-step record
static_assert(1 + 2 * kSplitDepth <= 16, "named barriers: 0 = __syncthreads, then full[D], empty[D]");

struct SplitCarve {
    int view_stride;
    size_t group_bytes, bars, group0, objs_off, ring_off;
    __host__ __device__ SplitCarve(int S, int SC, int TW, int TWD) {  // TWD = worlds of a group's tile (<= 32)
        view_stride = (int)align16((size_t)TWD * SC);
        objs_off = (size_t)2 * view_stride;
        ring_off = objs_off + align16((size_t)S * 32 * 2);
        group_bytes = ring_off + (size_t)kSplitDepth * kSplitWords * 32 * 4;
        bars = align16(sizeof(Tables)) + align16((size_t)SC);
        group0 = bars;
    }
};

template <int GE, int TW>
__global__ void __launch_bounds__(TW*(1 + GE) * 32) oc_rollout_split_kernel(const RolloutParams prm) {
    constexpr int P = 2, WPE = 32 / GE, WARPS = 1 + GE, D = kSplitDepth;
    extern __shared__ __align__(16) uint8_t smem[];
    stage_tables(smem, prm);
    const Tables& tb = *reinterpret_cast<const Tables*>(smem);
    const int SC = tb.SC, S = tb.S;
    const uint8_t* tmpl = smem + align16(sizeof(Tables));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp / WARPS, role = warp % WARPS;
    const int TWD = (prm.tile_worlds > 0 && prm.tile_worlds < 32) ? prm.tile_worlds : 32;  // worlds of a group's tile
    const SplitCarve cv(S, SC, TW, TWD);
    static_assert(TW == 1, "one group per CTA: the named barriers 1..2D are the group's");
    constexpr int full0 = 1, empty0 = 1 + D, kGroupThreads = WARPS * 32;

    const int view_stride = cv.view_stride;
    uint8_t* gbase = smem + cv.group0 + (size_t)group * cv.group_bytes;
    uint8_t* planes = gbase;                                               // [P][TWD][SC]
    uint16_t* objs = reinterpret_cast<uint16_t*>(gbase + cv.objs_off);     // [S][32], one column per world
    uint32_t* ring = reinterpret_cast<uint32_t*>(gbase + cv.ring_off);     // [D][kSplitWords][32]
    const int N = prm.N, K = prm.K;
    const int n0 = (blockIdx.x * TW + group) * TWD;
    if (n0 >= N) return;  // whole group idle (no block-level sync below)
    const int tile_valid = min(TWD, N - n0);
    const size_t PN = (size_t)P * N;
    const Consts c = load_consts(tb);

    if (role == 0) {
        // ================================================================ transition warp
        const int n = n0 + lane;
        const bool valid = lane < tile_valid;
        const int nl = valid ? n : N - 1;
        uint16_t* myobjs = objs + lane;
        World<P> w;
        load_world<P, 1, 8>(tb, c, prm, nl, 0, myobjs, w);
        int cur_return = prm.cur_return[nl];
        long long ret_add = 0;
        int ep_add = 0;
        const bool use_rng = prm.actions == nullptr;
        ActionRng<P> rng;
        unsigned long long t = prm.step0_dev != nullptr ? *prm.step0_dev : prm.step0;
        const uint32_t gworld = prm.world0 + (uint32_t)nl;
        if (use_rng && (t % ActionRng<P>::kStepsPerBlock) != 0) rng.refill(prm.seed, gworld, t);
        int32_t* rew_ptr = prm.rew ? prm.rew + n : nullptr;
        int32_t* done_ptr = prm.done ? prm.done + n : nullptr;
        uint8_t* aout_ptr = prm.actions_out ? prm.actions_out + n : nullptr;
        size_t act_idx = nl;
        const int pot0_cell = info_cell(c.pot0), pot1_cell = info_cell(c.pot1);

        for (int k = 0; k < K; ++k, ++t) {
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
                    for (int i = 0; i < P; ++i) aout_ptr[(size_t)i * N] = (uint8_t)act[i];
                }
                aout_ptr += PN;
            }
            uint32_t dirty[P], ticked;
            const int r = step_world<P>(tb, c, w, myobjs, 32, act, dirty, ticked);
            const bool done = w.timestep >= c.horizon;  // envs/overcooked2_env.py:334
            cur_return += r;
            if (done) {  // auto-reset, pantheonrl_extension/vectorenv.py:369-370
                ret_add += cur_return;
                ep_add += 1;
                cur_return = 0;
                reset_world<P>(tb, w);
                for (int idx = 0; idx < c.n_objcells; ++idx) myobjs[(int)tb.objcells[idx] * 32] = 0;
            }
            if (rew_ptr != nullptr) {
                if (valid) {
#pragma unroll
                    for (int i = 0; i < P; ++i) rew_ptr[(size_t)i * N] = r;
                }
                rew_ptr += PN;
            }
            if (done_ptr != nullptr) {
                if (valid) *done_ptr = done ? 1 : 0;
                done_ptr += N;
            }
            // ---- the record of this world-step (objects as they are AFTER the step, like obs_phase2 reads them)
            const int slot = k % D;
            if (k >= D) named_bar_sync(empty0 + slot, kGroupThreads);
            uint32_t* rec = ring + slot * (kSplitWords * 32) + lane;
            rec[0] = player_pack(w.pos[0], w.orient[0], w.held[0]) | (done ? (1u << 14) : 0u);
            rec[32] = player_pack(w.pos[1], w.orient[1], w.held[1]);
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const uint32_t d = dirty[i];
                const int cell = d == 0xFFFFFFFFu ? 0 : info_cell(d);
                rec[(2 + i) * 32] = d == 0xFFFFFFFFu ? 0xFFFFFFFFu : ((uint32_t)cell | ((uint32_t)myobjs[cell * 32] << 16));
            }
            uint32_t pots = 0u;
            if (ticked & 1u) pots |= (uint32_t)myobjs[pot0_cell * 32];
            if (ticked & 2u) pots |= (uint32_t)myobjs[pot1_cell * 32] << 16;
            rec[4 * 32] = pots;
            named_bar_arrive(full0 + slot, kGroupThreads);
        }
        if (prm.step_counter != nullptr && blockIdx.x == 0 && group == 0 && lane == 0)
            *prm.step_counter += (unsigned long long)K;
        if (valid) {
#pragma unroll
            for (int i = 0; i < P; ++i) prm.players[(size_t)i * N + n] = player_pack(w.pos[i], w.orient[i], w.held[i]);
            store_world_objs<1>(tb, c, prm, n, 0, myobjs);
            prm.timestep[n] = w.timestep;
            prm.cur_return[n] = cur_return;
            if (ep_add) {
                prm.ret_sum[n] += ret_add;
                prm.episodes[n] += ep_add;
            }
        }
        return;
    }

    // ==================================================================== encoder warps
    const int e = role - 1;
    const int wi = lane / GE, g = lane % GE;
    const int wt = e * WPE + wi;  // world inside the tile
    const int nl = min(n0 + wt, N - 1);
    const int c0 = n0 + e * WPE;  // first world of this warp's part of the tile
    const int nvalid = max(0, min(WPE, tile_valid - e * WPE));
    const bool lane_on = wt < tile_valid;  // the planes hold TWD worlds: lanes past the tile must not touch them
    uint8_t* myplanes = planes + (size_t)wt * SC;  // + v*view_stride
    const uint32_t* t4 = reinterpret_cast<const uint32_t*>(tmpl);
    const int n4 = SC >> 2;  // SC = 20*S

    // planes of the launch's initial state: template + the objects lying on counters / in pots (the players are poked
    // per step; the cells they stand on are AIR, all-zero in the template, and cleared again by the first step)
    int oldslot[P];
    if (nvalid > 0) {
        if (lane_on) {
#pragma unroll
            for (int v = 0; v < P; ++v) {
                uint32_t* d4 = reinterpret_cast<uint32_t*>(myplanes + v * view_stride);
                for (int j = g; j < n4; j += GE) d4[j] = t4[j];
            }
        }
        __syncwarp();
        constexpr int kB = 4;
        for (int i0 = g; lane_on && i0 < c.n_objcells; i0 += kB * GE) {
            int cell[kB];
            uint32_t o[kB];
#pragma unroll
            for (int j = 0; j < kB; ++j) {
                const int idx = i0 + j * GE;
                cell[j] = idx < c.n_objcells ? (int)tb.objcells[idx] : -1;
                o[j] = cell[j] >= 0 ? (uint32_t)prm.objs[(size_t)cell[j] * N + nl] : 0u;
            }
#pragma unroll
            for (int j = 0; j < kB; ++j) {
                if (o[j] == 0u) continue;
                const uint32_t ci = tb.cell_info[cell[j]];
#pragma unroll
                for (int v = 0; v < P; ++v) encode_cell<P>(myplanes + v * view_stride, ci, o[j]);
            }
        }
#pragma unroll
        for (int i = 0; i < P; ++i) oldslot[i] = info_slot(tb.cell_info[prm.players[(size_t)i * N + nl] & 0xFFFu]);
        __syncwarp();
    }

    const size_t obs_view_stride = (size_t)N * SC, obs_step_stride = PN * SC;
    int8_t* obs_ptr = prm.obs + (size_t)c0 * SC;
    const int nbytes = nvalid * SC;
    uint8_t* chunk = planes + (size_t)e * WPE * SC;  // + v*view_stride
    const uint32_t chunk_s = smem_u32(chunk);
    const bool tma_ok = prm.use_tma && ((nbytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(obs_ptr) & 15u) == 0) &&
                        ((obs_view_stride & 15u) == 0);
    bool pending = false;

    for (int k = 0; k < K; ++k) {
        const int slot = k % D;
        if (pending) {  // the previous bulk store must have finished reading the planes
            if (lane == 0) bulk_wait_read_all();
            pending = false;
        }
        named_bar_sync(full0 + slot, kGroupThreads);
        const uint32_t* rec = ring + slot * (kSplitWords * 32) + wt;
        const uint32_t pw[P] = {rec[0], rec[32]};
        const uint32_t dw[P] = {rec[64], rec[96]};
        const uint32_t pots = rec[128];
        if (k + D < K) named_bar_arrive(empty0 + slot, kGroupThreads);
        if (nvalid == 0) continue;
        const bool done = (pw[0] >> 14) & 1u;
        int slot_[P], orient[P];
        uint32_t held[P];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            slot_[i] = info_slot(tb.cell_info[pw[i] & 0xFFFu]);
            orient[i] = (int)((pw[i] >> 12) & 3u);
            held[i] = pw[i] >> 16;
        }
        // phase 1: rebuild from the template at an episode end, else clear the cells the players stood on
        if (!lane_on) {
        } else if (done) {
#pragma unroll
            for (int v = 0; v < P; ++v) {
                uint32_t* d4 = reinterpret_cast<uint32_t*>(myplanes + v * view_stride);
                for (int j = g; j < n4; j += GE) d4[j] = t4[j];
            }
        } else {
#pragma unroll
            for (int j0 = 0; j0 < P * P; j0 += GE) {
                const int j = j0 + g;
                if (j < P * P) clear_cell<P>(myplanes + (j / P) * view_stride, sel<P>(oldslot, j % P));
            }
        }
        __syncwarp();
        // phase 2: touched counter / pot cells, ticking pots, players
        if (lane_on && !done) {
#pragma unroll
            for (int j0 = 0; j0 < P * P; j0 += GE) {
                const int j = j0 + g;
                if (j < P * P) {
                    const uint32_t d = selu<P>(dw, j / P);
                    if (d != 0xFFFFFFFFu) {
                        const uint32_t ci = tb.cell_info[d & 0xFFu];
                        uint32_t b0, w14;
                        cell_bytes(ci, d >> 16, b0, w14);
                        store_cell<P>(myplanes + (j % P) * view_stride, ci, b0, w14);
                    }
                }
            }
            for (int j = g; j < c.n_pots * P; j += GE) {
                const int q = j / P;
                const uint32_t o = q ? (pots >> 16) : (pots & 0xFFFFu);
                if (o != 0u) {
                    const uint32_t ci = q ? c.pot1 : c.pot0;
                    uint32_t b0, w14;
                    cell_bytes(ci, o, b0, w14);
                    store_cell<P>(myplanes + (j % P) * view_stride, ci, b0, w14);
                }
            }
        }
#pragma unroll
        for (int j0 = 0; j0 < P * P; j0 += GE) {
            const int j = j0 + g;
            if (lane_on && j < P * P) {
                const int v = j / P, i = j % P;
                poke_player<P>(myplanes + v * view_stride, v, i, sel<P>(slot_, i), sel<P>(orient, i), selu<P>(held, i));
            }
        }
#pragma unroll
        for (int i = 0; i < P; ++i) oldslot[i] = slot_[i];
        if (tma_ok) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int v = 0; v < P; ++v)
                    bulk_store_s2g(obs_ptr + v * obs_view_stride, chunk_s + (uint32_t)(v * view_stride), (uint32_t)nbytes);
                bulk_commit();
            }
            pending = true;
        } else {
            __syncwarp();
#pragma unroll
            for (int v = 0; v < P; ++v) warp_copy_out(obs_ptr + v * obs_view_stride, chunk + v * view_stride, nbytes, lane);
        }
        obs_ptr += obs_step_stride;
    }
    if (pending && lane == 0) bulk_wait_read_all();
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
    const int TWD = (!observe_only && prm.tile_worlds > 0 && prm.tile_worlds < WPW) ? prm.tile_worlds : WPW;
    const int tiles = (prm.N + TWD - 1) / TWD;
    const int ctas = (tiles + warps_per_cta - 1) / warps_per_cta;
    auto kern = observe_only ? oc_observe_kernel<P, G> : (prm.K <= 2 ? oc_rollout_kernel<P, G, true> : oc_rollout_kernel<P, G, false>);
    if (smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
    }
    kern<<<ctas, warps_per_cta * 32, smem_bytes, stream>>>(prm);
    return cudaGetLastError();
}

size_t rollout_split_smem_bytes(int S, int C, int TW, int tile_worlds) {
    const SplitCarve cv(S, S * C, TW, (tile_worlds > 0 && tile_worlds < 32) ? tile_worlds : 32);
    return cv.group0 + (size_t)TW * cv.group_bytes;
}

template <int GE, int TW>
static cudaError_t launch_split_t(const RolloutParams& prm, size_t smem_bytes, cudaStream_t stream) {
    auto kern = oc_rollout_split_kernel<GE, TW>;
    if (smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
    }
    const int TWD = (prm.tile_worlds > 0 && prm.tile_worlds < 32) ? prm.tile_worlds : 32;
    const int groups = (prm.N + TWD - 1) / TWD;
    kern<<<(groups + TW - 1) / TW, TW*(1 + GE) * 32, smem_bytes, stream>>>(prm);
    return cudaGetLastError();
}

cudaError_t launch_rollout_split(const RolloutParams& prm, int GE, int TW, size_t smem_bytes, cudaStream_t stream) {
    if (GE == 4 && TW == 1) return launch_split_t<4, 1>(prm, smem_bytes, stream);
    if (GE == 2 && TW == 1) return launch_split_t<2, 1>(prm, smem_bytes, stream);
    return cudaErrorInvalidValue;
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
