// oc_device.cuh — device-only helpers shared by oc_kernels.cu and the fused policy rollout
// (policy_kernels.cu): bulk (TMA) stores, streaming copy-out, state load.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "oc_core.cuh"
#include "oc_kernels.h"

namespace ocb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_s2g(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- rebuild of a warp tile's planes by ONE bulk copy per view (global template tile -> shared), tracked by an mbarrier.
// The per-lane rebuild (obs_phase1 with full = true) costs SC / 4 shared stores per world and view; at one env step per
// launch that loop was the largest single item of the kernel (ncu source view, 16 % of the instructions).
__device__ __forceinline__ void tile_fill_begin(uint32_t bar, uint32_t planes_s, int view_stride, const uint8_t* tile,
                                                uint32_t bytes, int views) {  // one lane
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * (uint32_t)views) : "memory");
    for (int v = 0; v < views; ++v)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         planes_s + (uint32_t)(v * view_stride)),
                     "l"(tile), "r"(bytes), "r"(bar)
                     : "memory");
}
__device__ __forceinline__ void tile_fill_wait(uint32_t bar) {  // every lane that goes on to touch the planes
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(0u)
            : "memory");
        if (ok) return;
        if (spins > (1u << 24)) __trap();
    }
}

// ---- named-barrier hand-offs (bar.arrive by the producer, bar.sync by the consumer; whole warps, `threads` = all
// participants of both sides).  A waiting warp is descheduled by the hardware: unlike an mbarrier poll loop it takes no
// issue slots from the warps that are working, which is what the split rollout kernel's transition warp needs.
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ int load_action(const void* a, int dtype, size_t idx) {
    int v;
    switch (dtype) {
        case OCB_ACT_I64: v = (int)static_cast<const long long*>(a)[idx]; break;
        case OCB_ACT_F32: v = (int)static_cast<const float*>(a)[idx]; break;
        case OCB_ACT_U8: v = (int)static_cast<const uint8_t*>(a)[idx]; break;
        default: v = static_cast<const int*>(a)[idx]; break;
    }
    return (v >= 0 && v <= 5) ? v : A_STAY;
}

// streams `nbytes` of shared memory to global memory with the widest aligned stores
__device__ __forceinline__ void warp_copy_out(int8_t* __restrict__ dst, const uint8_t* __restrict__ src, int nbytes,
                                              int lane) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        const int n16 = nbytes >> 4;
        const uint4* s16 = reinterpret_cast<const uint4*>(src);
        uint4* d16 = reinterpret_cast<uint4*>(dst);
        for (int c = lane; c < n16; c += 32) __stcs(d16 + c, s16[c]);
        for (int b = (n16 << 4) + lane; b < nbytes; b += 32) dst[b] = (int8_t)src[b];
    } else if ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
        const int n4 = nbytes >> 2;
        const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
        for (int c = lane; c < n4; c += 32) __stcs(d4 + c, s4[c]);
        for (int b = (n4 << 2) + lane; b < nbytes; b += 32) dst[b] = (int8_t)src[b];
    } else {
        for (int b = lane; b < nbytes; b += 32) dst[b] = (int8_t)src[b];
    }
}

// HBM -> registers / shared memory
template <int P, int G, int kBatch = 8>
__device__ __forceinline__ void load_world(const Tables& tb, const Consts& c, const RolloutParams& prm, int nl, int g,
                                           uint16_t* myobjs, World<P>& w) {
    const int N = prm.N;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const uint32_t pw = prm.players[(size_t)i * N + nl];
        w.pos[i] = (int)(pw & 0xFFFu);
        w.slot[i] = info_slot(tb.cell_info[w.pos[i]]);
        w.orient[i] = (int)((pw >> 12) & 3u);
        w.held[i] = pw >> 16;
    }
    w.timestep = prm.timestep[nl];
    // `myobjs` is this LANE's private column of the warp's [S][32] object array: the G lanes of a world split the
    // global loads and write each value into all G sibling columns (disjoint cells per writer), after which every
    // lane only ever touches its own column.  Only counters and pots can hold an object (step_world, import check):
    // the other cells are neither loaded nor stored, nor ever read from shared memory.
    uint16_t* col0 = myobjs - g;  // column of the world's lane 0
    int cd = 0, np = 0;
    // batches of independent loads: one at a time (a dependent chain objcells[idx] -> address -> value -> store) left a
    // single-step launch waiting ~700 cycles per cell (a third of its stall samples, ncu source view)
    for (int i0 = g; i0 < c.n_objcells; i0 += kBatch * G) {
        int cell[kBatch];
        uint32_t o[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const int idx = i0 + j * G;
            cell[j] = idx < c.n_objcells ? (int)tb.objcells[idx] : -1;
            o[j] = cell[j] >= 0 ? (uint32_t)prm.objs[(size_t)cell[j] * N + nl] : 0u;
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            if (cell[j] < 0) continue;
#pragma unroll
            for (int h = 0; h < G; ++h) col0[cell[j] * 32 + h] = (uint16_t)o[j];
            const int t = info_terrain(tb.cell_info[cell[j]]);
            cd += (t == T_COUNTER && obj_name(o[j]) == O_DISH);
            np += (t == T_POT) ? pot_counts(o[j]) : 0;
        }
    }
    __syncwarp();
#pragma unroll
    for (int m = 1; m < G; m <<= 1) {
        cd += __shfl_xor_sync(0xffffffffu, cd, m);
        np += __shfl_xor_sync(0xffffffffu, np, m);
    }
    w.counter_dishes = cd;
    w.nonempty_pots = np;
}

// shared memory -> HBM: the cells that can hold an object (the counterpart of load_world)
template <int G>
__device__ __forceinline__ void store_world_objs(const Tables& tb, const Consts& c, const RolloutParams& prm, int n, int g,
                                                 const uint16_t* myobjs) {
    for (int idx = g; idx < c.n_objcells; idx += G) {
        const int cell = (int)tb.objcells[idx];
        prm.objs[(size_t)cell * prm.N + n] = myobjs[cell * 32];
    }
}

}  // namespace ocb
