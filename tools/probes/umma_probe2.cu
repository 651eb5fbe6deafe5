// umma_probe2.cu — round-2 extension of umma_probe.cu: kind::tf32, M=64 and the 128B-swizzle SS mode next to the
// TS mode the policy kernels use (MODE 0 = SS no-swizzle, 1 = SS 128B swizzle, 2 = TS: A from TMEM, B no-swizzle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu && ./umma_probe
// Reports cycles per MMA (M=128, K=16, bf16, SS mode, no-swizzle K-major operands) as a function of
// N, of the number of accumulators the MMAs rotate over, and of the number of distinct A tiles.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int tf32) {
    const uint32_t fmt = tf32 ? 2u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_ts_tf32(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

template <int MODE, int TF>
__global__ void __launch_bounds__(128, 1) probe(int M, int N, int nacc, int nA, int iters, int swz, int commit_every, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_addr(&bar2)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const uint32_t a0 = smem_addr(smem), b0 = smem_addr(smem + 128 * 1024);
        const uint32_t idesc = make_idesc(M, N, TF);
        const int CE = 0;
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 3; ++rep) {
            t0 = clock64();
            if (elect_one()) {
                // descriptors advance by masks only: the loop must not be issue-bound by integer division
                const uint64_t da0 = MODE == 1 ? (make_desc(a0, 16, 1024) | (2ull << 61)) : make_desc(a0, 128, 256);
                const uint64_t db0 = MODE == 1 ? (make_desc(b0, 16, 1024) | (2ull << 61)) : make_desc(b0, 128, 256);
                const uint32_t astep = MODE == 1 ? 2 : 256;
                const uint32_t bar2a = smem_addr(&bar2);
#pragma unroll 8
                for (int i = 0; i < iters; ++i) {
                    const uint32_t d = tmem + (i & (nacc - 1)) * N;
                    if (MODE == 2) {
                        if (TF) umma_ts_tf32(d, tmem + 448 + (i & (nA - 1)) * 8, db0, idesc, i >= nacc);
                        else umma_ts(d, tmem + 448 + (i & (nA - 1)) * 8, db0, idesc, i >= nacc);
                    } else {
                        if (TF) umma_tf32(d, da0 + (uint64_t)((i & (nA - 1)) * astep), db0, idesc, i >= nacc);
                        else umma(d, da0 + (uint64_t)((i & (nA - 1)) * astep), db0, idesc, i >= nacc);
                    }
                    if (CE && (i & (CE - 1)) == CE - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar2a) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar)) : "memory");
            }
            __syncwarp();
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(ok)
                             : "r"(smem_addr(&bar)), "r"((uint32_t)(rep & 1))
                             : "memory");
            }
            t1 = clock64();
        }
        if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int MODE, int TF>
static double run(int M, int N, int nacc, int nA, long long* d_out) {
    const int iters = 512;
    cudaFuncSetAttribute(probe<MODE, TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe<MODE, TF><<<148, 128, 200 * 1024>>>(M, N, nacc, nA, iters, 0, 0, d_out);
    long long c = 0;
    cudaError_t e = cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        printf("error: %s\n", cudaGetErrorString(e));
        exit(1);
    }
    return (double)c / iters;
}
int main() {
    long long* d_out;
    cudaMalloc(&d_out, 8);
    printf("%4s %6s %5s %-22s %-5s %12s\n", "M", "N", "nacc", "mode", "kind", "cyc/MMA");
    for (int M : {128, 64})
        for (int N : {32, 64})
            for (int nacc : {1, 2}) {
                printf("%4d %6d %5d %-22s %-5s %12.1f\n", M, N, nacc, "SS no-swizzle", "bf16", run<0, 0>(M, N, nacc, 8, d_out));
                printf("%4d %6d %5d %-22s %-5s %12.1f\n", M, N, nacc, "SS 128B-swizzle", "bf16", run<1, 0>(M, N, nacc, 8, d_out));
                printf("%4d %6d %5d %-22s %-5s %12.1f\n", M, N, nacc, "TS (B no-swizzle)", "bf16", run<2, 0>(M, N, nacc, 8, d_out));
                printf("%4d %6d %5d %-22s %-5s %12.1f\n", M, N, nacc, "SS no-swizzle", "tf32", run<0, 1>(M, N, nacc, 8, d_out));
                printf("%4d %6d %5d %-22s %-5s %12.1f\n", M, N, nacc, "SS 128B-swizzle", "tf32", run<1, 1>(M, N, nacc, 8, d_out));
                printf("%4d %6d %5d %-22s %-5s %12.1f\n", M, N, nacc, "TS (B no-swizzle)", "tf32", run<2, 1>(M, N, nacc, 8, d_out));
            }
    return 0;
}
