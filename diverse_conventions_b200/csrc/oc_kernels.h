// oc_kernels.h — host-visible launch interface of oc_kernels.cu / bb_kernels.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ocb {

struct Tables;

constexpr int kThreadsPerCta = 128;

__host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

struct RolloutParams {
    const Tables* tables;  // device copy of the static layout tables
    const uint8_t* tmpl;   // device, [32][S*C]: the static part of one observation plane, repeated for the 32 worlds of a
                           // warp tile (row 0 is what the per-world rebuild reads; the whole tile is the source of the bulk
                           // copy that rebuilds a warp's planes at the start of a launch)
    // world state, structure-of-arrays in HBM
    uint32_t* players;   // [P][N]  pos | orient<<12 | held<<16
    uint16_t* objs;      // [S][N]  packed object per cell
    int32_t* timestep;   // [N]
    int32_t* cur_return; // [N] return of the running episode
    long long* ret_sum;  // [N] sum of returns of completed episodes
    int32_t* episodes;   // [N] completed episodes
    int N;
    int K;
    unsigned long long step0;  // global step counter of the first step of this launch (host mirror)
    const unsigned long long* step0_dev;  // when set, the launch reads its first step from the device counter instead
                                          // (launches captured into a CUDA graph: replays must not reuse step0)
    unsigned long long seed;
    unsigned int world0;  // global index of world 0 (multi-GPU shards draw disjoint streams)
    const void* actions;  // [K][P][N] or nullptr -> on-device RNG
    int act_dtype;
    uint8_t* actions_out;  // [K][P][N] or nullptr
    int8_t* obs;           // [K][P][N][W][H][C] or nullptr
    int32_t* rew;          // [K][P][N] or nullptr
    int32_t* done;         // [K][N] or nullptr
    int use_tma;
    int tile_worlds;  // worlds per warp tile (one-warp kernel: <= 32 / G) resp. per group (split kernel: <= 32); 0 = full
                      // tiles.  Narrower tiles spread a launch evenly over all SMs (see ocb_api.cu: balanced_tile)
    unsigned long long* step_counter;  // device mirror of the global step counter (+= K per launch) or nullptr
};

size_t rollout_smem_bytes(int P, int S, int C, int G, int warps_per_cta);

cudaError_t launch_rollout(const RolloutParams& prm, int P, int G, int warps_per_cta, size_t smem_bytes,
                           bool observe_only, cudaStream_t stream);
// role-split K-step kernel (P = 2, <= 2 pots, observations wanted): GE encoder warps per transition warp, TW groups per CTA
size_t rollout_split_smem_bytes(int S, int C, int TW, int tile_worlds);
cudaError_t launch_rollout_split(const RolloutParams& prm, int GE, int TW, size_t smem_bytes, cudaStream_t stream);
cudaError_t launch_counter_add(unsigned long long* counter, unsigned long long k, cudaStream_t stream);
cudaError_t launch_reset(const Tables* tables, uint32_t* players, uint16_t* objs, int32_t* timestep, int32_t* cur_return,
                         int N, int rows, cudaStream_t stream);
cudaError_t launch_export_state(const Tables* tables, const uint32_t* players, const uint16_t* objs,
                                const int32_t* timestep, int32_t* out, int N, cudaStream_t stream);
cudaError_t launch_import_state(const Tables* tables, const int32_t* in, uint32_t* players, uint16_t* objs,
                                int32_t* timestep, int32_t* cur_return, int N, int* bad, cudaStream_t stream);

}  // namespace ocb
