// policy_generic.cuh — the MAPPO actor / critic forward for every shape the tensor-core kernels do not take
// (included by policy_kernels.cu inside its anonymous namespace).
//
// The tcgen05 kernels (policy_kernels.cu, policy512.cuh) are specialised for 2 players (20 observation channels packed
// into 16 bf16 slots) and grids up to 6 rows high (a grid column per warp-wide load, 192 TMEM columns of cells).  The
// reference's CNNBase (train/MAPPO/utils/cnn.py:22-42) takes any (W, H, C): schelling (7 x 7), corridor (14 x 9),
// multiplayer_schelling (7 x 7, 4 players = 30 channels) and simple_single (1 player = 15 channels) train there.  This
// kernel covers those: same network, same entry points, fp32 FMAs on the CUDA cores — one CTA per tile of kGRows rows,
// the conv and the FC1 partial sum fused position by position so that no [rows x npos x h/2] activation is ever
// materialised:
//   for every conv position p:  a[r][co] = relu(b[co] + sum_{i,j,c} w[co][c][i][j] * obs[r][ox+i][oy+j][c])   (staged in smem)
//                               acc1[r][n] += sum_co a[r][co] * W1[n][co * npos + p]                          (registers)
//   h1 = relu(acc1 + b1);  acc2 = h1 x W2^T;  h2 = relu(acc2 + b2);  head = h2 x Wh^T + bh  -> logits | value
// Weights are fp32, transposed on the host so that the threads of a warp (consecutive output units) read consecutive
// words; they are served from L2.  Sampling / log-prob / entropy / stored-action evaluation go through emit_actor_row,
// exactly as in the tensor-core kernels (same counters, same arithmetic).  Accuracy: fp32 accumulation in a different
// order than torch (~1e-6 relative).  Bound: FMA issue; this is a coverage path, measured in profiles/README.md.
#pragma once

constexpr int kGRows = 16;      // rows per CTA tile (divides the 128-row tile_policy granularity)
constexpr int kGenThreads = 256;

// per-network blob (floats): conv_wt [9*C][CO] | conv_b [CO] | w1t [npos][CO][h] | b1 [h] | w2t [h][h] | b2 [h] |
// wh [8][h] | bh [8]
struct GBlob {
    size_t conv_w, conv_b, w1, b1, w2, b2, wh, bh, total;  // offsets in floats
};
__host__ __device__ inline GBlob gblob_layout(int C, int hidden, int npos) {
    GBlob L;
    const size_t CO = hidden / 2;
    size_t o = 0;
    L.conv_w = o, o += (size_t)9 * C * CO;
    L.conv_b = o, o += CO;
    L.w1 = o, o += (size_t)npos * CO * hidden;
    L.b1 = o, o += hidden;
    L.w2 = o, o += (size_t)hidden * hidden;
    L.b2 = o, o += hidden;
    L.wh = o, o += (size_t)8 * hidden;
    L.bh = o, o += 8;
    L.total = (o + 31) & ~(size_t)31;
    return L;
}

struct GParams {
    PolicyParams base;    // obs, M, outputs, sampling parameters, tile_policy, row_index, given_actions ...
    const float* blobs;   // [n_policies][2 nets][GBlob]
    int C, hidden;
};

__host__ __device__ inline size_t generic_smem_bytes(int SC, int hidden) {
    // obs bytes of the tile | conv activations [rows][CO] | two activation planes [rows][hidden] | head partials
    size_t o = ((size_t)kGRows * SC + 15) & ~(size_t)15;
    o += (size_t)kGRows * (hidden / 2) * 4;
    o += 2 * (size_t)kGRows * hidden * 4;
    return o + 128;
}

// kH = hidden size (64 / 512).  Thread t owns output units n = t, t + 256, ... (kH / 256 of them, or one unit shared by
// 256 / kH row groups when kH < 256) for a subset of the rows.
template <int kH>
__global__ void __launch_bounds__(kGenThreads) policy_generic_kernel(const GParams gp) {
    constexpr int CO = kH / 2;
    constexpr int kUnits = kH >= kGenThreads ? kH / kGenThreads : 1;          // units per thread
    constexpr int kRowGroups = kH >= kGenThreads ? 1 : kGenThreads / kH;      // row groups when one thread = one unit
    constexpr int kRowsPerThread = kGRows / kRowGroups;
    const PolicyParams& prm = gp.base;
    extern __shared__ __align__(16) uint8_t gsm[];
    const int W = prm.W, H = prm.H, C = gp.C, SC = prm.SC, npos = prm.npos, PH = H - 2;
    int8_t* s_obs = reinterpret_cast<int8_t*>(gsm);
    float* s_conv = reinterpret_cast<float*>(gsm + (((size_t)kGRows * SC + 15) & ~(size_t)15));  // [rows][CO]
    float* s_h1 = s_conv + kGRows * CO;                                                            // [rows][kH]
    float* s_h2 = s_h1 + kGRows * kH;                                                              // [rows][kH]
    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kGRows;
    const int nrows = (int)min((long long)kGRows, (long long)prm.M - row0);
    const GBlob L = gblob_layout(C, kH, npos);
    const int pol = prm.tile_policy ? prm.tile_policy[row0 / 128] : 0;
    unsigned long long offset = prm.offset;
    if (prm.d_offset != nullptr) offset += *prm.d_offset;

    // observations of the tile (rows past M re-read the last row; their results are never stored)
    if ((SC & 3) == 0 && (reinterpret_cast<uintptr_t>(prm.obs) & 3u) == 0) {
        const int words = SC >> 2;
        const uint32_t* obs32 = reinterpret_cast<const uint32_t*>(prm.obs);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_obs);
        for (int i = tid; i < kGRows * words; i += kGenThreads) {
            const int r = i / words, w = i - r * words;
            long long row = row0 + (r < nrows ? r : nrows - 1);
            if (prm.row_index != nullptr) row = prm.row_index[row];
            dst[i] = __ldg(obs32 + row * words + w);
        }
    } else {  // 4 players on an odd-sized grid: 30 S bytes per row, not a multiple of 4
        for (int i = tid; i < kGRows * SC; i += kGenThreads) {
            const int r = i / SC, b = i - r * SC;
            long long row = row0 + (r < nrows ? r : nrows - 1);
            if (prm.row_index != nullptr) row = prm.row_index[row];
            s_obs[i] = __ldg(prm.obs + row * SC + b);
        }
    }
    __syncthreads();

    const int my_unit0 = kH >= kGenThreads ? tid : tid % kH;
    const int my_r0 = kH >= kGenThreads ? 0 : (tid / kH) * kRowsPerThread;

    for (int net = 0; net < 2; ++net) {
        if (!((prm.net_mask >> net) & 1)) continue;
        const float* blob = gp.blobs + ((size_t)pol * 2 + net) * L.total;
        float acc[kUnits][kRowsPerThread];
#pragma unroll
        for (int q = 0; q < kUnits; ++q)
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) acc[q][r] = 0.0f;

        for (int p = 0; p < npos; ++p) {
            const int ox = p / PH, oy = p - ox * PH;
            // conv position p for all rows: kGRows * CO outputs over the block
            for (int o = tid; o < kGRows * CO; o += kGenThreads) {
                const int r = o / CO, co = o - r * CO;
                float a = __ldg(blob + L.conv_b + co);
                const int8_t* orow = s_obs + (size_t)r * SC;
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) {
                        const int8_t* cell = orow + ((ox + i) * H + (oy + j)) * C;
                        const float* wk = blob + L.conv_w + (size_t)((i * 3 + j) * C) * CO + co;
                        for (int c = 0; c < C; ++c) a = fmaf((float)cell[c], __ldg(wk + (size_t)c * CO), a);
                    }
                s_conv[o] = fmaxf(a, 0.0f);
            }
            __syncthreads();
            // FC1 partial sum of this position
            const float* w1p = blob + L.w1 + (size_t)p * CO * kH;
            for (int co = 0; co < CO; ++co) {
                float wv[kUnits];
#pragma unroll
                for (int q = 0; q < kUnits; ++q) wv[q] = __ldg(w1p + (size_t)co * kH + my_unit0 + q * kGenThreads);
#pragma unroll
                for (int r = 0; r < kRowsPerThread; ++r) {
                    const float a = s_conv[(my_r0 + r) * CO + co];
#pragma unroll
                    for (int q = 0; q < kUnits; ++q) acc[q][r] = fmaf(a, wv[q], acc[q][r]);
                }
            }
            __syncthreads();
        }
        // h1 = relu(acc1 + b1)
#pragma unroll
        for (int q = 0; q < kUnits; ++q) {
            const int n = my_unit0 + q * kGenThreads;
            const float b = __ldg(blob + L.b1 + n);
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) s_h1[(my_r0 + r) * kH + n] = fmaxf(acc[q][r] + b, 0.0f);
        }
        __syncthreads();
        // FC2
#pragma unroll
        for (int q = 0; q < kUnits; ++q)
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) acc[q][r] = 0.0f;
        for (int k = 0; k < kH; ++k) {
            float wv[kUnits];
#pragma unroll
            for (int q = 0; q < kUnits; ++q) wv[q] = __ldg(blob + L.w2 + (size_t)k * kH + my_unit0 + q * kGenThreads);
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) {
                const float a = s_h1[(my_r0 + r) * kH + k];
#pragma unroll
                for (int q = 0; q < kUnits; ++q) acc[q][r] = fmaf(a, wv[q], acc[q][r]);
            }
        }
#pragma unroll
        for (int q = 0; q < kUnits; ++q) {
            const int n = my_unit0 + q * kGenThreads;
            const float b = __ldg(blob + L.b2 + n);
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) s_h2[(my_r0 + r) * kH + n] = fmaxf(acc[q][r] + b, 0.0f);
        }
        __syncthreads();
        // head: one warp per row (8 warps take the 16 rows in two rounds), shuffle reduction over the hidden units
        const int warp = tid >> 5, lane = tid & 31;
        for (int r = warp; r < kGRows; r += kGenThreads / 32) {
            float hv[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            const int nout = net == 0 ? 6 : 1;
            for (int k = lane; k < kH; k += 32) {
                const float h = s_h2[r * kH + k];
                for (int a = 0; a < nout; ++a) hv[a] = fmaf(h, __ldg(blob + L.wh + (size_t)a * kH + k), hv[a]);
            }
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) hv[a] += __shfl_xor_sync(0xffffffffu, hv[a], s);
            if (lane == 0 && r < nrows) {
                float head[6];
#pragma unroll
                for (int a = 0; a < 6; ++a) head[a] = hv[a] + (a < nout ? __ldg(blob + L.bh + a) : 0.0f);
                const long long row = row0 + r;
                if (net == 1) {
                    if (prm.values) prm.values[row] = head[0];
                } else {
                    emit_actor_row(prm, row, (uint32_t)row, head, offset);
                }
            }
        }
        __syncthreads();
    }
}
