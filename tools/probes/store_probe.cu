// store_probe.cu — how fast can ONE SM push bytes to HBM, and what is the write-only ceiling of the GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/store_probe tools/probes/store_probe.cu
// Every CTA streams `iters` tiles of `tile` bytes from shared memory to its own region of a large global buffer, with
// bulk stores (mode 0: one elected lane per warp, one bulk store per warp-tile) or 16-byte st.global.cs (mode 1).
// Output: one line per (grid, warps, tile, mode): aggregate GB/s and GB/s per CTA.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void store_kernel(uint8_t* dst, size_t region, int tile, int iters, int mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    uint8_t* mine = smem + (size_t)warp * tile;
    for (int i = lane * 4; i < tile; i += 128) *reinterpret_cast<uint32_t*>(mine + i) = i ^ blockIdx.x;
    __syncwarp();
    uint8_t* out = dst + (size_t)blockIdx.x * region + (size_t)warp * tile;
    const size_t stride = (size_t)nw * tile;
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(mine);
    for (int it = 0; it < iters; ++it) {
        if (mode >= 2) {  // bulk store with an L2 eviction-priority hint (2 evict_first, 3 evict_last, 4 evict_unchanged)
            if (lane == 0) {
                uint64_t pol;
                if (mode == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
                else if (mode == 3) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
                else asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(out), "r"(s),
                             "r"(tile), "l"(pol)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            __syncwarp();
        } else if (mode == 0) {
            if (lane == 0) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out), "r"(s), "r"(tile) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            __syncwarp();
        } else {
            const uint4* s16 = reinterpret_cast<const uint4*>(mine);
            uint4* d16 = reinterpret_cast<uint4*>(out);
            for (int c = lane; c < (tile >> 4); c += 32) __stcs(d16 + c, s16[c]);
        }
        out += stride;
    }
    if (mode != 1 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
    const size_t total = (size_t)6 << 30;
    uint8_t* d;
    cudaMalloc(&d, total);
    cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const bool hints = argc > 1;  // any argument: the L2-hint modes on the oversubscribed shapes only
    const int grids_a[] = {1, 16, 64, 128, 148, 296, 592}, grids_b[] = {128, 148, 296, 592};
    const int warps_a[] = {1, 4, 8}, warps_b[] = {4};
    const int tiles_a[] = {3200, 12800, 25600}, tiles_b[] = {12800};
    const int* grids = hints ? grids_b : grids_a; const int ng = hints ? 4 : 7;
    const int* warpss = hints ? warps_b : warps_a; const int nwp = hints ? 1 : 3;
    const int* tiles = hints ? tiles_b : tiles_a; const int nt = hints ? 1 : 3;
    const char* names[] = {"bulk", "stg128", "bulk_evict_first", "bulk_evict_last", "bulk_evict_unchanged"};
    for (int mode = hints ? 0 : 0; mode < (hints ? 5 : 2); ++mode)
        for (int ti = 0; ti < nt; ++ti)
            for (int wi = 0; wi < nwp; ++wi)
                for (int gi = 0; gi < ng; ++gi) {
                    const int tile = tiles[ti], warps = warpss[wi], grid = grids[gi];
                    if ((size_t)warps * tile > 200 * 1024) continue;
                    const size_t region = total / grid / 128 * 128;
                    int iters = (int)(region / ((size_t)warps * tile));
                    const size_t want = ((size_t)3 << 30) / grid;  // ~3 GB per launch, less for tiny grids
                    if ((size_t)iters * warps * tile > want) iters = (int)(want / ((size_t)warps * tile));
                    if (grid < 64 && iters > 4000) iters = 4000;
                    float best = 1e30f;
                    for (int rep = 0; rep < 3; ++rep) {
                        cudaEventRecord(e0);
                        store_kernel<<<grid, warps * 32, warps * tile>>>(d, region, tile, iters, mode);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                        float ms;
                        cudaEventElapsedTime(&ms, e0, e1);
                        if (ms < best) best = ms;
                    }
                    const double bytes = (double)grid * iters * warps * tile;
                    printf("{\"mode\": \"%s\", \"tile\": %d, \"warps\": %d, \"grid\": %d, \"GBs\": %.1f, \"GBs_per_cta\": %.2f, \"err\": \"%s\"}\n",
                           names[mode], tile, warps, grid, bytes / best / 1e6, bytes / best / 1e6 / grid,
                           cudaGetErrorString(cudaGetLastError()));
                }
    return 0;
}
