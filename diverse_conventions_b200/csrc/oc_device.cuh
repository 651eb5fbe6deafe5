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
template <int P, int G>
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
    // lane only ever touches its own column
    uint16_t* col0 = myobjs - g;  // column of the world's lane 0
    for (int cell = g; cell < tb.S; cell += G) {
        const uint16_t v = prm.objs[(size_t)cell * N + nl];
#pragma unroll
        for (int h = 0; h < G; ++h) col0[cell * 32 + h] = v;
    }
    __syncwarp();
    int cd = 0, np = 0;
    for (int idx = g; idx < c.n_objcells; idx += G) {
        const uint32_t ci = tb.cell_info[tb.objcells[idx]];
        const uint32_t o = myobjs[info_cell(ci) * 32];
        cd += (info_terrain(ci) == T_COUNTER && obj_name(o) == O_DISH);
        np += (info_terrain(ci) == T_POT) ? pot_counts(o) : 0;
    }
#pragma unroll
    for (int m = 1; m < G; m <<= 1) {
        cd += __shfl_xor_sync(0xffffffffu, cd, m);
        np += __shfl_xor_sync(0xffffffffu, np, m);
    }
    w.counter_dishes = cd;
    w.nonempty_pots = np;
}

}  // namespace ocb
