// policy512.cuh — MAPPO actor / critic forward for hidden_size 512 (the argparse default,
// train/config.py:199; every train/*.sh uses 64, which policy_kernels.cu fuses into one kernel).
// Textually included inside the anonymous namespace of policy_kernels.cu, after the shared helpers.
//
//   obs int8 [M, W, H, 20] -> Conv3x3(20->256) -> ReLU -> FC(256 npos -> 512) -> ReLU -> FC(512 -> 512) -> ReLU -> head
//
// At this width the layers are GEMM-sized (2.66 MFLOP per row and network), the activations of a
// 128-row tile no longer fit on one SM (conv output 128 x 1536 for cramped_room), so the forward is
// three persistent tcgen05 kernels per launch that hand the activations over in HBM / L2:
//   conv512_kernel : loader_role (shared with the h=64 kernels) -> per-cell bf16 blocks in TMEM -> per
//                    position two N=128 accumulates (A from TMEM, B = resident conv weights, 147 KB of shared
//                    memory) -> bias/ReLU -> hi/lo split -> stores in the packed operand layout of FC1;
//   gemm512_kernel : C[128 x 512] = A[128 x K] W^T; A and W arrive as pre-packed [rows x 16] canonical
//                    K-major blocks (hi | lo), 8 KB + 32 KB of cp.async.bulk per K step into a 4-stage
//                    ring, 6 tcgen05.mma (SS mode, M128 N256 K16; 3 products hi.hi + hi.lo + lo.hi) per step,
//                    the whole 512-column accumulator in TMEM; epilogue = bias/ReLU/split -> packed stores
//                    (FC1) or the head, sampling and outputs (FC2).
// Every operand block is stored in HBM exactly as the MMA reads it from shared memory (8-row x 16-byte
// core matrices, no swizzle), so all copies are 1-D bulk copies and the epilogue stores are full 128-byte
// segments.  Precision scheme as for h=64: bf16 hi + lo operands, fp32 accumulate (~1e-5 relative).

constexpr int kH5 = 512;                       // hidden size
constexpr int kCo5 = 256;                      // conv output channels
constexpr int kCw5 = kCo5 * kK1 * 2;           // 73,728 B: conv weights [256 x 144] bf16 canonical, hi (or lo)
// K extent of one packed block = one MMA step.  (32 until round 2: with 80 KB per ring stage only two stages fit, and FC1 —
// whose K loop is 48-126 blocks long — ran at 5.5 TB/s of L2 reads where the three-stage FC2 reached 8.9: it was bound by the
// depth of its ring, not by L2 bandwidth.  Halving the block doubles the stages in the same shared memory.)
constexpr int kKc = 16;
constexpr int kKSbo = kKc / 8 * 128;           // bytes between 8-row groups of a packed block
constexpr int kABlk = 2 * kRows * kKc * 2;     // 8 KB: activations [128 x 16] bf16 canonical, hi | lo
constexpr int kWBlk = 2 * kH5 * kKc * 2;       // 32 KB: weights [512 x 16] bf16 canonical, hi | lo
constexpr int kKc2 = kH5 / kKc;                // 32 K blocks of FC2
constexpr int kGStages = 5;                    // operand ring of gemm512_kernel (one pass: 5 x 40 KB; two passes: six stages of 24 KB)
constexpr int kGWarpMma = kEpiWarps, kGWarpProd = kEpiWarps + 1;
constexpr int kGThreads = 32 * (kEpiWarps + 2);  // 320
constexpr int kC5ColD1 = 192;                  // conv accumulators: half h at + 128 h

// packed weight blob of one network (device): all offsets are multiples of 128
struct Blob5 {
    size_t wc_hi, wc_lo, bias1, w1, b1, w2, b2, wh, bh, total;
    int kc1;  // K blocks of FC1 = 256 npos / kKc
};
__host__ __device__ inline Blob5 blob5_layout(int npos) {
    Blob5 L;
    size_t o = 0;
    L.kc1 = kCo5 * npos / kKc;
    L.wc_hi = o, o += kCw5;
    L.wc_lo = o, o += kCw5;
    L.bias1 = o, o += (size_t)npos * kCo5 * 4;
    L.w1 = o, o += (size_t)L.kc1 * kWBlk;
    L.b1 = o, o += kH5 * 4;
    L.w2 = o, o += (size_t)kKc2 * kWBlk;
    L.b2 = o, o += kH5 * 4;
    L.wh = o, o += 8 * kH5 * 4;
    L.bh = o, o += 128;
    L.total = o;
    return L;
}

struct P5Params {
    PolicyParams base;  // geometry, observations, outputs, unit mapping (blobs / blob_stride address Blob5 blobs)
    uint8_t* a1;        // conv output = FC1 operand: [2 nets][tiles][kc1][kABlk]
    uint8_t* a2;        // FC1 output = FC2 operand:  [2 nets][tiles][kKc2][kABlk]
    int kc1;
};

// A and B from shared memory (SS mode)
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

// thread `r` (tile row) stores 32 consecutive K values (hi/lo packed pairs) into the two consecutive blocks at `blk` of the
// packed operand layout: per block [r/8][k8][r%8][8 x bf16] (hi half, then lo half); a warp writes full 128-byte segments
__device__ __forceinline__ void store_packed32(uint8_t* blk, int r, const uint32_t (&hi)[16], const uint32_t (&lo)[16]) {
    static_assert(kKc == 16, "two blocks of 16 per call");
    uint8_t* p = blk + (r >> 3) * kKSbo + (r & 7) * 16;
#pragma unroll
    for (int k8 = 0; k8 < 4; ++k8) {
        uint8_t* d = p + (k8 >> 1) * kABlk + (k8 & 1) * 128;
        *reinterpret_cast<uint4*>(d) = make_uint4(hi[4 * k8], hi[4 * k8 + 1], hi[4 * k8 + 2], hi[4 * k8 + 3]);
        *reinterpret_cast<uint4*>(d + kABlk / 2) = make_uint4(lo[4 * k8], lo[4 * k8 + 1], lo[4 * k8 + 2], lo[4 * k8 + 3]);
    }
}

// ---------------------------------------------------------------- conv512_kernel
enum : int {
    C5_COL_FULL = B_COL_FULL,    // [4] shared with loader_role
    C5_COL_EMPTY = B_COL_EMPTY,  // [4]
    C5_D1_FULL = 8,              // [2] MMA commit -> epilogue group h
    C5_D1_EMPTY = 10,            // [2] epilogue group h (128 arrivals) -> MMA
    C5_W_FULL = 12,              //     bulk copies -> MMA, epilogue
    C5_W_EMPTY = 13,             // [2] by unit parity: MMA commit + 256 epilogue arrivals -> producer
    C5_COUNT = 15,
    C5_TMEM_SLOT = 24
};
struct C5Smem {
    int stage, wc, bias1, bars, total;
};
__host__ __device__ inline C5Smem c5_smem_layout(int npos, int stage_stride) {
    C5Smem s;
    int o = 0;
    s.stage = o, o += al128(kLoadWarps * 32 * stage_stride * 4);
    s.wc = o, o += 2 * kCw5;
    s.bias1 = o, o += npos * kCo5 * 4;
    s.bars = o, o += 256;
    s.total = o + 128;
    return s;
}

__global__ void __launch_bounds__(kThreads, 1) conv512_kernel(const P5Params q) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_addr(smem_raw) & 127u)) & 127u);
    const PolicyParams& prm = q.base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int npos = prm.npos;
    const Blob5 L = blob5_layout(npos);
    const C5Smem sl = c5_smem_layout(npos, prm.stage_stride);
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(smem + sl.stage);
    uint8_t* s_wc = smem + sl.wc;
    const float* s_bias1 = reinterpret_cast<const float*>(smem + sl.bias1);
    uint64_t* s_bars = reinterpret_cast<uint64_t*>(smem + sl.bars);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bars + C5_TMEM_SLOT);
    const uint32_t bars = smem_addr(s_bars);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)),
                     "r"((uint32_t)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < C5_COUNT; ++i) {
            uint32_t count = 1;
            if ((i >= C5_COL_FULL && i < C5_COL_FULL + 4) || (i >= C5_D1_EMPTY && i < C5_D1_EMPTY + 2)) count = 128;
            if (i >= C5_W_EMPTY && i < C5_W_EMPTY + 2) count = 32 * kEpiWarps + 1;
            mbar_init(bars + 8 * i, count);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const UnitRange ur = my_units(prm);
    long long pw[1] = {};

    if (warp < kEpiWarps) {
        // ---- epilogue: group g drains accumulator half g (conv channels 128 g .. 128 g + 127)
        const int g = warp >> 2, r = (warp & 3) * 32 + (tid & 31);
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t u = 0, gen = 0, cnt = 0;
        for (int t = ur.t0; t < ur.t1; ++t, ++u) {
            if (blob_changed(prm, t, ur.t0)) {
                mbar_wait(bars + 8 * C5_W_FULL, gen & 1);
                ++gen;
            }
            uint8_t* blk0 = q.a1 + ((size_t)ur.net * prm.tiles + t) * q.kc1 * kABlk;
            for (int p = 0; p < npos; ++p, ++cnt) {
                mbar_wait(bars + 8 * (C5_D1_FULL + g), cnt & 1);
                tc_fence_after();
#pragma unroll 1
                for (int b = 0; b < 4; ++b) {
                    float v[32];
                    tmem_ld32(trow + kC5ColD1 + g * 128 + b * 32, v);
                    if (b == 3) {  // the half is in registers: the MMA may overwrite it with the next position
                        tc_fence_before();
                        mbar_arrive(bars + 8 * (C5_D1_EMPTY + g));
                    }
                    const float* bias = s_bias1 + p * kCo5 + g * 128 + b * 32;
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        split2(fmaxf(v[2 * i] + bias[2 * i], 0.0f), fmaxf(v[2 * i + 1] + bias[2 * i + 1], 0.0f), hi[i], lo[i]);
                    store_packed32(blk0 + (size_t)(p * 8 + g * 4 + b) * 2 * kABlk, r, hi, lo);
                }
            }
            mbar_arrive(bars + 8 * (C5_W_EMPTY + (u & 1)));  // done with bias1 of this unit
        }
    } else if (warp < kWarpMma) {
        loader_role<false>(pw, prm, ur, tmem, s_stage, bars);
    } else if (warp == kWarpMma) {
        // ---- MMA issuer
        const int W = prm.W, H = prm.H, PH = H - 2;
        const uint32_t idesc = make_idesc(kRows, 128);
        const uint32_t a_wchi = smem_addr(s_wc), a_wclo = a_wchi + kCw5;
        uint32_t gcb = 0, u = 0, gen = 0, cnt = 0;
        for (int t = ur.t0; t < ur.t1; ++t, ++u, gcb += W) {
            if (blob_changed(prm, t, ur.t0)) {
                mbar_wait(bars + 8 * C5_W_FULL, gen & 1);
                ++gen;
            }
            int ox = 0, oy = 0;
            for (int p = 0; p < npos; ++p, ++cnt) {
                if (oy == 0) {  // new window column(s)
                    for (int d = (ox == 0 ? 0 : 2); d < 3; ++d) {
                        const uint32_t gcol = gcb + ox + d;
                        mbar_wait(bars + 8 * (C5_COL_FULL + gcol % kColRing), (gcol / kColRing) & 1);
                    }
                }
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    if (cnt > 0) mbar_wait(bars + 8 * (C5_D1_EMPTY + h), (cnt - 1) & 1);
                    tc_fence_after();
                    const uint32_t d1 = tmem + kC5ColD1 + h * 128;
                    const uint32_t boff = (uint32_t)h * (128 / 8) * 2304;  // 16 row groups of the canonical [256 x 144] operand
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < 9; ++j) {
                            const int dx = j / 3, dy = j - dx * 3;
                            const uint32_t ta = tmem + kColCells + (((gcb + ox + dx) % kColRing) * H + oy + dy) * kCellCols;
                            umma_bf16_ts(d1, ta, make_desc(a_wchi + boff + j * 256, 128, 2304), idesc, j > 0);
                            umma_bf16_ts(d1, ta, make_desc(a_wclo + boff + j * 256, 128, 2304), idesc, 1);
                        }
                        umma_commit(bars + 8 * (C5_D1_FULL + h));
                        if (h == 1) {
                            if (oy == PH - 1) {  // the window leaves column ox (and the last two columns with the last window)
                                umma_commit(bars + 8 * (C5_COL_EMPTY + (gcb + ox) % kColRing));
                                if (ox == W - 3) {
                                    umma_commit(bars + 8 * (C5_COL_EMPTY + (gcb + ox + 1) % kColRing));
                                    umma_commit(bars + 8 * (C5_COL_EMPTY + (gcb + ox + 2) % kColRing));
                                }
                            }
                            if (p + 1 == npos) umma_commit(bars + 8 * (C5_W_EMPTY + (u & 1)));
                        }
                    }
                    __syncwarp();
                }
                if (++oy == PH) oy = 0, ++ox;
            }
        }
    } else {
        // ---- producer: conv weights + per-position bias, resident until a tile selects another policy
        uint32_t u = 0;
        const uint32_t s_dst = smem_addr(s_wc);
        for (int t = ur.t0; t < ur.t1; ++t, ++u) {
            if (!blob_changed(prm, t, ur.t0)) continue;
            if (u > 0) mbar_wait(bars + 8 * (C5_W_EMPTY + ((u - 1) & 1)), ((u - 1) >> 1) & 1);
            if (elect_one()) {
                const uint8_t* blob = prm.blobs + ((size_t)tile_pol(prm, t) * 2 + ur.net) * prm.blob_stride;
                const uint32_t bias_bytes = (uint32_t)npos * kCo5 * 4;
                const uint32_t wb = bars + 8 * C5_W_FULL;
                mbar_arrive_expect_tx(wb, 2u * kCw5 + bias_bytes);
                constexpr int kPiece = kCw5 / 4;  // 18,432 B
#pragma unroll 1
                for (int i = 0; i < 8; ++i) bulk_g2s(s_dst + i * kPiece, blob + L.wc_hi + (size_t)i * kPiece, kPiece, wb);
                bulk_g2s(s_dst + 2 * kCw5, blob + L.bias1, bias_bytes, wb);
            }
            __syncwarp();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
}

// ---------------------------------------------------------------- gemm512_kernel
// Barrier slots (8 bytes each): operand ring FULL / EMPTY [3 each], accumulator FULL / EMPTY per N half [2 each]
enum : int { G5_FULL = 0, G5_EMPTY = 6, G5_D_FULL = 12, G5_D_EMPTY = 14, G5_COUNT = 16, G5_TMEM_SLOT = 20 };
struct G5Smem {
    int stages, xbuf, bias, wh, bars, total;
};
__host__ __device__ inline G5Smem g5_smem_layout() {
    G5Smem s;
    int o = 0;
    s.stages = o, o += kGStages * (kABlk + kWBlk);
    s.xbuf = o, o += kRows * 8 * 4;
    s.bias = o, o += kH5 * 4;
    s.wh = o, o += 8 * kH5 * 4 + 128;  // head weights (FC2: hidden-major [512][8], see below) + head bias
    s.bars = o, o += 256;
    s.total = o + 128;
    return s;
}

// kHead = false: FC1 (A = q.a1, K blocks = kc1, W = w1, bias b1) -> q.a2 in the packed layout.
// kHead = true : FC2 (A = q.a2, 32 K blocks, W = w2, bias b2) -> head, sampling, outputs.
// kTwo  = false: one pass over K with the whole 512-column accumulator; both epilogue groups drain it afterwards (FC1).
// kTwo  = true : TWO passes over K, one per N half (FC2, round 2): the head makes that epilogue long (512 x 6 FMAs per row),
//                and with one pass the tensor pipe idled under it (26 % of the burst peak against FC1's 62 %).  Now epilogue
//                group h drains half h while the K loop of the other half — or of the next tile's half — runs; the A block
//                is fetched twice (8 KB of the 24 KB per step), and the ring is six stages deep in the same shared memory.  Measured: hidden-512 forward of cramped_room 0.492 -> 0.405 ms per 32,768 rows.  FC1 streams
//                three times the weights and sits near the L2 bound: two passes cost it 2-4 %, it keeps one.
template <bool kHead, bool kTwo>
__global__ void __launch_bounds__(kGThreads, 1) gemm512_kernel(const P5Params q) {
    constexpr int kPasses = kTwo ? 2 : 1;                           // N halves computed one after the other
    constexpr int kWPass = kWBlk / kPasses;                         // weight bytes per K step and pass (hi | lo)
    constexpr int kStageBytes = kABlk + kWPass;
    constexpr int kStages = kTwo ? 6 : kGStages;
    static_assert(kStages * kStageBytes <= kGStages * (kABlk + kWBlk), "operand ring exceeds its shared memory");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_addr(smem_raw) & 127u)) & 127u);
    const PolicyParams& prm = q.base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const Blob5 L = blob5_layout(prm.npos);
    const G5Smem sl = g5_smem_layout();
    float* s_xbuf = reinterpret_cast<float*>(smem + sl.xbuf);
    float* s_bias = reinterpret_cast<float*>(smem + sl.bias);
    float* s_wh = reinterpret_cast<float*>(smem + sl.wh);
    uint64_t* s_bars = reinterpret_cast<uint64_t*>(smem + sl.bars);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bars + G5_TMEM_SLOT);
    const uint32_t bars = smem_addr(s_bars);
    const uint32_t s_stage0 = smem_addr(smem + sl.stages);
    const int KC = kHead ? kKc2 : q.kc1;
    const size_t w_off = kHead ? L.w2 : L.w1;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)),
                     "r"((uint32_t)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        // accumulator EMPTY: one pass -> both groups (256 threads) on slot 0; two passes -> group h (128 threads) on slot h
        for (int i = 0; i < G5_COUNT; ++i)
            mbar_init(bars + 8 * i, (i >= G5_D_EMPTY && i < G5_D_EMPTY + 2) ? (kTwo ? 128 : 32 * kEpiWarps) : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const UnitRange ur = my_units(prm);

    if (warp < kEpiWarps) {
        // ---- epilogue: group g drains accumulator columns 256 g .. 256 g + 255
        const int g = warp >> 2, r = (warp & 3) * 32 + (tid & 31), et = tid;  // et: 0..255
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        unsigned long long offset = prm.offset;
        if (kHead && prm.d_offset != nullptr) offset += *prm.d_offset;
        uint32_t u = 0;
        for (int t = ur.t0; t < ur.t1; ++t, ++u) {
            if (blob_changed(prm, t, ur.t0)) {  // stage this network's bias (and head) while the K loop runs
                const uint8_t* blob = prm.blobs + ((size_t)tile_pol(prm, t) * 2 + ur.net) * prm.blob_stride;
                asm volatile("bar.sync 2, 256;" ::: "memory");  // everybody is done with the previous set
                const float* gb = reinterpret_cast<const float*>(blob + (kHead ? L.b2 : L.b1));
                for (int i = et; i < kH5; i += 256) s_bias[i] = gb[i];
                if (kHead) {
                    // blob: wh [8][512] (a-major) then bh [8].  Staged hidden-major [512][8] for the actor (two 16-byte
                    // broadcast loads per hidden unit instead of six scalar ones), as it is ([512]) for the critic.
                    const float* gw = reinterpret_cast<const float*>(blob + L.wh);
                    if (ur.net == 0) {
                        for (int i = et; i < 8 * kH5; i += 256) s_wh[(i & (kH5 - 1)) * 8 + (i >> 9)] = gw[i];
                    } else {
                        for (int i = et; i < kH5; i += 256) s_wh[i] = gw[i];
                    }
                    for (int i = et; i < 8; i += 256) s_wh[8 * kH5 + i] = gw[8 * kH5 + i];
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            mbar_wait(bars + 8 * (G5_D_FULL + (kTwo ? g : 0)), u & 1);
            tc_fence_after();
            float head[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            uint8_t* blk0 = q.a2 + ((size_t)ur.net * prm.tiles + t) * kKc2 * kABlk;
#pragma unroll 1
            for (int b = 0; b < 8; ++b) {
                float v[32];
                tmem_ld32(trow + g * 256 + b * 32, v);
                if (b == 7) {  // the accumulator (half) is in registers: the MMA may overwrite it
                    tc_fence_before();
                    mbar_arrive(bars + 8 * (G5_D_EMPTY + (kTwo ? g : 0)));
                }
                const int n0 = g * 256 + b * 32;
                if (!kHead) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        split2(fmaxf(v[2 * i] + s_bias[n0 + 2 * i], 0.0f), fmaxf(v[2 * i + 1] + s_bias[n0 + 2 * i + 1], 0.0f), hi[i], lo[i]);
                    store_packed32(blk0 + (size_t)(n0 / kKc) * kABlk, r, hi, lo);  // n0 is a multiple of 32: two blocks
                } else if (ur.net == 0) {
                    const float4* w4 = reinterpret_cast<const float4*>(s_wh) + 2 * n0;
                    const float4* b4 = reinterpret_cast<const float4*>(s_bias + n0);
#pragma unroll
                    for (int qd = 0; qd < 8; ++qd) {
                        const float4 bb = b4[qd];
                        const float bq[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int i = 4 * qd + k;
                            const float h = fmaxf(v[i] + bq[k], 0.0f);
                            const float4 wa = w4[2 * i], wb = w4[2 * i + 1];
                            head[0] = fmaf(h, wa.x, head[0]), head[1] = fmaf(h, wa.y, head[1]), head[2] = fmaf(h, wa.z, head[2]);
                            head[3] = fmaf(h, wa.w, head[3]), head[4] = fmaf(h, wb.x, head[4]), head[5] = fmaf(h, wb.y, head[5]);
                        }
                    }
                } else {
                    const float4* w4 = reinterpret_cast<const float4*>(s_wh + n0);
                    const float4* b4 = reinterpret_cast<const float4*>(s_bias + n0);
#pragma unroll
                    for (int qd = 0; qd < 8; ++qd) {
                        const float4 bb = b4[qd], w = w4[qd];
                        head[0] = fmaf(fmaxf(v[4 * qd + 0] + bb.x, 0.0f), w.x, head[0]);
                        head[0] = fmaf(fmaxf(v[4 * qd + 1] + bb.y, 0.0f), w.y, head[0]);
                        head[0] = fmaf(fmaxf(v[4 * qd + 2] + bb.z, 0.0f), w.z, head[0]);
                        head[0] = fmaf(fmaxf(v[4 * qd + 3] + bb.w, 0.0f), w.w, head[0]);
                    }
                }
            }
            if (kHead) {
                // group 1 hands its partial sums to group 0 (same rows: warps w and w + 4)
                float* xrow = s_xbuf + (size_t)r * 8;
                if (g == 1) {
                    *reinterpret_cast<float4*>(xrow) = make_float4(head[0], head[1], head[2], head[3]);
                    *reinterpret_cast<float2*>(xrow + 4) = make_float2(head[4], head[5]);
                }
                asm volatile("bar.sync %0, 64;" ::"r"(3 + (warp & 3)) : "memory");
                if (g == 0) {
                    const float4 x0 = *reinterpret_cast<const float4*>(xrow);
                    const float2 x1 = *reinterpret_cast<const float2*>(xrow + 4);
                    const float* bh = s_wh + 8 * kH5;
                    head[0] += x0.x + bh[0], head[1] += x0.y + bh[1], head[2] += x0.z + bh[2], head[3] += x0.w + bh[3];
                    head[4] += x1.x + bh[4], head[5] += x1.y + bh[5];
                    const long long row = (long long)t * kRows + r;
                    if (row < prm.M) {
                        if (ur.net == 1) {
                            if (prm.values) prm.values[row] = head[0];
                        } else {
                            emit_actor_row(prm, row, (uint32_t)row, head, offset);
                        }
                    }
                }
                asm volatile("bar.sync %0, 64;" ::"r"(3 + (warp & 3)) : "memory");  // xbuf is rewritten by the next unit
            }
        }
    } else if (warp == kGWarpMma) {
        // ---- MMA issuer
        const uint32_t idesc = make_idesc(kRows, 256);
        uint32_t it = 0, u = 0;
        for (int t = ur.t0; t < ur.t1; ++t, ++u) {
#pragma unroll 1
            for (int pass = 0; pass < kPasses; ++pass) {
                // the accumulator columns of this pass have been drained by their epilogue group(s)
                if (u > 0) mbar_wait(bars + 8 * (G5_D_EMPTY + pass), (u - 1) & 1);
                for (int kc = 0; kc < KC; ++kc, ++it) {
                    const uint32_t s = it % kStages;
                    mbar_wait(bars + 8 * (G5_FULL + s), (it / kStages) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = s_stage0 + s * kStageBytes, a_lo = a_hi + kABlk / 2;
                    const uint32_t b_hi = a_hi + kABlk, b_lo = b_hi + kWPass / 2;
                    if (elect_one()) {
                        const uint64_t dahi = make_desc(a_hi, 128, kKSbo), dalo = make_desc(a_lo, 128, kKSbo);
#pragma unroll
                        for (int nh = 0; nh < 2 / kPasses; ++nh) {  // one pass: both N halves from one weight block
                            const uint32_t dst = tmem + (kPasses == 2 ? pass : nh) * 256;
                            const uint32_t boff = (uint32_t)nh * (256 / 8) * kKSbo;  // 32 row groups per N half
                            const uint64_t dbhi = make_desc(b_hi + boff, 128, kKSbo), dblo = make_desc(b_lo + boff, 128, kKSbo);
                            umma_bf16_ss(dst, dahi, dbhi, idesc, kc != 0 ? 1u : 0u);
                            umma_bf16_ss(dst, dahi, dblo, idesc, 1);
                            umma_bf16_ss(dst, dalo, dbhi, idesc, 1);
                        }
                        umma_commit(bars + 8 * (G5_EMPTY + s));
                        if (kc + 1 == KC) umma_commit(bars + 8 * (G5_D_FULL + pass));
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ---- producer: one activation block and one weight block (one pass: 64 KB; two passes: the pass's N half, hi and lo)
        const uint8_t* A = kHead ? q.a2 : q.a1;
        uint32_t it = 0;
        for (int t = ur.t0; t < ur.t1; ++t) {
            const uint8_t* blob = prm.blobs + ((size_t)tile_pol(prm, t) * 2 + ur.net) * prm.blob_stride;
            const uint8_t* a_src = A + ((size_t)ur.net * prm.tiles + t) * KC * kABlk;
#pragma unroll 1
            for (int pass = 0; pass < kPasses; ++pass) {
                for (int kc = 0; kc < KC; ++kc, ++it) {
                    const uint32_t s = it % kStages;
                    if (it >= kStages) mbar_wait(bars + 8 * (G5_EMPTY + s), ((it / kStages) - 1) & 1);
                    if (elect_one()) {
                        const uint32_t fb = bars + 8 * (G5_FULL + s);
                        const uint32_t dst = s_stage0 + s * kStageBytes;
                        mbar_arrive_expect_tx(fb, kStageBytes);
                        bulk_g2s(dst, a_src + (size_t)kc * kABlk, kABlk, fb);
                        const uint8_t* w_src = blob + w_off + (size_t)kc * kWBlk;
                        if (kPasses == 2) {  // rows 256 pass .. of the hi block and of the lo block (16 KB each, contiguous)
                            bulk_g2s(dst + kABlk, w_src + (size_t)pass * (kWBlk / 4), kWBlk / 4, fb);
                            bulk_g2s(dst + kABlk + kWBlk / 4, w_src + kWBlk / 2 + (size_t)pass * (kWBlk / 4), kWBlk / 4, fb);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) bulk_g2s(dst + kABlk + i * (kWBlk / 4), w_src + (size_t)i * (kWBlk / 4), kWBlk / 4, fb);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
}
