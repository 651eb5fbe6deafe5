// umma_probe3.cu — why do the conv MMAs of the policy kernels run at ~60 cycles each when umma_probe2 measures 32.5
// for the same shape (M128 N64 K16, TS mode)?  One factor at a time, one CTA per SM:
//   bmode 0: every MMA reads the SAME B tile (as umma_probe2)
//   bmode 1: B rotates over 9 K-slices of a [64 x 144] no-swizzle operand (LBO 128, SBO 2304: the conv weights)
//   bmode 2: as 1, alternating between two operands 18 KB apart (hi / lo)
//   bmode 3: B rotates over 8 [64 x 32] chunks (LBO 128, SBO 512: the FC weights)
//   dol    : the {0,0,0,0} disable-output-lane form of the instruction (what policy_kernels.cu issues)
//   traffic: 0 none | 1 eight other warps loop tcgen05.ld 32x32b.x32 | 2 they loop tcgen05.st x16 + wait |
//            3 they stream shared memory with 16-byte loads (the B operand competes for the 128 B/clk pipe) |
//            4 they loop ld x32 + 2 x st x16 like the epilogue groups |
//            5 they run ALU-bound loops (issue-slot competition on the issuer's SM sub-partition)
//   nA     : distinct A tiles in TMEM
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe3_probe umma_probe3.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_ts_dol(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(d),
        "r"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

template <int BMODE, int DOL>
__global__ void __launch_bounds__(384, 1) probe(int N, int nacc, int nA, int iters, int traffic, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) stop = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const uint32_t b0 = smem_addr(smem);
        const uint32_t idesc = make_idesc(128, N);
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 3; ++rep) {
            t0 = clock64();
            if (elect_one()) {
                uint64_t db[18];
#pragma unroll
                for (int j = 0; j < 18; ++j) {
                    if (BMODE == 0) db[j] = make_desc(b0, 128, 256);
                    if (BMODE == 1) db[j] = make_desc(b0 + (j % 9) * 256, 128, 2304);
                    if (BMODE == 2) db[j] = make_desc(b0 + (j & 1) * 18432 + (j >> 1) * 256, 128, 2304);
                    if (BMODE == 3) db[j] = make_desc(b0 + (j % 8) * 8192 + ((j / 8) & 1) * 256, 128, 512);
                }
                uint32_t da[18], dd[18];  // addresses precomputed: the loop below must be nothing but MMA issues
#pragma unroll
                for (int j = 0; j < 18; ++j) da[j] = tmem + 384 + (j % nA) * 8, dd[j] = tmem + (j & (nacc - 1)) * N;
                for (int i0 = 0; i0 < iters; i0 += 18) {
#pragma unroll
                    for (int j = 0; j < 18; ++j) {
                        if (DOL) umma_ts_dol(dd[j], da[j], db[j], idesc, 1u);
                        else umma_ts(dd[j], da[j], db[j], idesc, 1u);
                    }
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
        __syncwarp();
        if (threadIdx.x == 32) stop = 1;
    } else if (warp >= 4 && traffic != 0) {
        // TMEM traffic on columns 256..383 of this warp's lane quadrant
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + ((warp >> 2) & 1) * 64;
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = threadIdx.x + i;
        while (!stop) {
            if (traffic == 1) {
                uint32_t v[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                r[0] ^= v[3];
            } else if (traffic == 5) {
                float f0 = (float)r[0], f1 = (float)r[1], f2 = (float)r[2], f3 = (float)r[3];
#pragma unroll 16
                for (int i = 0; i < 256; ++i) {
                    f0 = fmaf(f0, 1.0001f, 0.5f), f1 = fmaf(f1, 0.9999f, 0.25f), f2 = fmaf(f2, 1.0002f, 0.125f), f3 = fmaf(f3, 0.9998f, 1.0f);
                }
                r[0] = __float_as_uint(f0 + f1 + f2 + f3);
            } else if (traffic == 3) {
                const uint4* sp = reinterpret_cast<const uint4*>(smem + 64 * 1024) + threadIdx.x;
                uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll 8
                for (int i = 0; i < 64; ++i) {
                    const uint4 q = sp[(i & 7) * 384];
                    acc.x ^= q.x, acc.y ^= q.y, acc.z ^= q.z, acc.w ^= q.w;
                }
                r[0] ^= acc.x ^ acc.y ^ acc.z ^ acc.w;
            } else {
                if (traffic == 4) {
                    uint32_t v[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] ^= v[i] + v[16 + i];
                    asm volatile(
                        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
                            "r"(taddr + 32), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                        : "memory");
                }
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
                        "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                    : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }
        if (r[0] == 0x12345678u) out[1] = r[0];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int BMODE, int DOL>
static double run(int N, int nacc, int nA, int traffic, long long* d_out) {
    const int iters = 18 * 32;
    cudaFuncSetAttribute(probe<BMODE, DOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe<BMODE, DOL><<<148, 384, 200 * 1024>>>(N, nacc, nA, iters, traffic, d_out);
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
    cudaMalloc(&d_out, 16);
    printf("%4s %5s %4s %-28s %4s %-8s %10s\n", "N", "nacc", "nA", "B operand", "dol", "traffic", "cyc/MMA");
    const char* tn[6] = {"none", "ld x32", "st x16", "smem", "epi-like", "alu"};
    for (int N : {64, 32})
        for (int traffic : {0, 5})
            for (int nA : {12}) {
                const int nacc = 2;
                printf("%4d %5d %4d %-28s %4d %-8s %10.1f\n", N, nacc, nA, "same tile", 0, tn[traffic], run<0, 0>(N, nacc, nA, traffic, d_out));
                printf("%4d %5d %4d %-28s %4d %-8s %10.1f\n", N, nacc, nA, "same tile", 1, tn[traffic], run<0, 1>(N, nacc, nA, traffic, d_out));
                printf("%4d %5d %4d %-28s %4d %-8s %10.1f\n", N, nacc, nA, "9 K-slices of [64x144]", 1, tn[traffic], run<1, 1>(N, nacc, nA, traffic, d_out));
                printf("%4d %5d %4d %-28s %4d %-8s %10.1f\n", N, nacc, nA, "hi/lo x 9 slices", 1, tn[traffic], run<2, 1>(N, nacc, nA, traffic, d_out));
                printf("%4d %5d %4d %-28s %4d %-8s %10.1f\n", N, nacc, nA, "hi/lo x 9 slices", 0, tn[traffic], run<2, 0>(N, nacc, nA, traffic, d_out));
                printf("%4d %5d %4d %-28s %4d %-8s %10.1f\n", N, nacc, nA, "8 FC chunks [64x32]", 1, tn[traffic], run<3, 1>(N, nacc, nA, traffic, d_out));
            }
    return 0;
}
