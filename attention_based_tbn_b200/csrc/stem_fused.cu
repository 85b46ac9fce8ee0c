// Fused network stem for cin = 1 (audio), cin = 3 (RGB) and cin = 10 (flow):
//     conv1_7x7_s2 + BN + ReLU  ->  pool1 (MaxPool 3x3 / 2, ceil_mode)  ->  conv2_3x3_reduce (1x1) + BN + ReLU
// (reference core/models/bn_inception_audio.py:21-31, 35-39, dataflow 417-426) in ONE tcgen05 kernel that reads the
// NCHW fp32 network input and writes the NHWC bf16 conv2_3x3_reduce output.  The conv1 output (1.6 MB per RGB frame,
// 3.5 MB per audio frame -- the largest activation of the network) and the pool1 output never exist in memory.
//
// A tile = up to 63 pooled pixels of ONE pooled row.  They need conv1 rows 2*py .. 2*py+2 over 128 conv columns, i.e.
// 11 input rows x 264 input columns per channel, staged once (coalesced) as fp16.  Per conv row every thread (= one
// conv pixel = one TMEM lane) gathers its operand row from the staged rows (chunk (c, r) = 8 consecutive staged values)
// and writes it with tcgen05.st straight into TENSOR MEMORY: the conv1 MMAs take A from TMEM, which (measured,
// scripts/exp/mma_rate.cu) costs ~40 cycles per N = 64 instruction instead of the ~62 the shared-memory operand path
// needs, and the operand tile is double buffered there, so building row q+1 overlaps the MMAs of row q.  One thread
// issues the MMAs (the folded conv1 bias rides on two extra weight columns), every thread reads its pixel's 64
// accumulators and folds them into a running packed-fp16 maximum that starts at zero (= ReLU + vertical max).  The
// horizontal 3-tap maximum goes through a swizzled shared-memory tile, which IS the K-major operand tile of the 1x1
// conv2_3x3_reduce GEMM (64 channels = one 128-byte row); its epilogue adds the bias, applies ReLU and one thread
// issues a TMA tensor store of the 63 x 64 result.  conv1 row 2*py+2 is recomputed by the next pooled row (1.5x the
// conv1 MMAs, which are ~2 % of the network's FLOPs) -- the price of not writing / re-reading 5 MB per frame.
//
// Operand precision: INSIDE this kernel the tensor-core operands are fp16, not bf16 (kind::f16 takes either at the same
// rate): the staged input, the conv1 weights, the conv1 output / pooled tile and the conv2_3x3_reduce weights.  The first
// layers difference large, smooth inputs (log-spectrogram values of -14 .. 10, images with a DC offset), so their
// rounding errors are amplified by cancellation: emulating the bf16 backbone on the CPU shows these four rounding
// points alone contribute ~1.2e-2 of the ~1.5-2.1e-2 logit error of an all-bf16 backbone, and
// fp16's three extra mantissa bits remove them for free.  The values are bounded (inputs |x| < 16, BatchNorm-folded
// activations), far inside fp16's range.  What leaves the kernel (the conv2_3x3_reduce output) is bf16 like every other
// activation of the backbone.
#include <cuda_fp16.h>

#include <type_traits>

#include "tc_common.cuh"

namespace tbn {

struct StemFusedParams {
    const void* x;       // (n, cin, h, w) fp32 or bf16 (staged as fp16: a bf16 input converts exactly)
    const float* bias2;  // (64) folded bias of conv2_3x3_reduce
    int n, h, w;         // network input
    int oh, ow;          // conv1 output
    int ph, pw;          // pool1 output = conv2_3x3_reduce output
    int segs;            // ceil(pw / 63) tiles per pooled row
    int tiles;           // n * ph * segs
    unsigned long long* trace;  // diagnostic (tbn_debug_trace_stem): clock64 stamps of CTA 0's first tiles, 16 per tile
};
#define SF_STAMP(i)                                                                                 \
    do {                                                                                            \
        if (p.trace != nullptr && blockIdx.x == 0 && tid == 0 && tcount < 8)                        \
            p.trace[tcount * 16 + (i)] = (unsigned long long)clock64();                             \
    } while (0)

// four 32-bit columns of this thread's TMEM lane (32x32b: lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v0), "r"(v1), "r"(v2), "r"(v3)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// kind::f16 instruction descriptor with fp16 A and B (format code 0; bf16 is 1), fp32 accumulation, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_f16(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A = 128 lanes x 8 columns (16 16-bit values of K per lane)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

constexpr int SF_SROW = 264;  // staged fp16 per (channel, input row): 2*128 + 8
constexpr int SF_ROWS = 11;   // input rows behind three conv rows (7 + 2 + 2)
constexpr int SF_RING = 12;   // staged row slots per channel: input row iy lives in slot iy mod 12, so that the 7 rows a
                              // tile shares with the pooled row below it (4 input rows further down) stay where they are
constexpr int SF_PW = 63;     // pooled pixels per tile (conv columns 2*px .. 2*px+2 inside 128 conv columns)
constexpr int SF_COUT = 64;
constexpr int SF_B_BYTES = SF_COUT * 128;

__device__ __forceinline__ float ld_in(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_in(const __nv_bfloat16* p) { return __bfloat162float(*p); }  // (fp32 branch only)
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// packed bf16 pair -> packed fp16 pair (exact for |x| < 65504 above fp16's subnormal range)
__device__ __forceinline__ uint32_t bf2_to_h2(uint32_t v) {
    return pack_h2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}

template <int CIN, typename TIN>
__global__ void __launch_bounds__(128, CIN == 1 ? 4 : (CIN == 3 ? 2 : 1)) stem_fused_kernel(const __grid_constant__ CUtensorMap tmW1,
                                                                          const __grid_constant__ CUtensorMap tmW2,
                                                                          const __grid_constant__ CUtensorMap tmY,
                                                                          const StemFusedParams p) {
    constexpr int NCH = CIN * 7;       // 16-byte operand chunks per conv pixel (one per (c, r))
    constexpr int KB = (NCH + 8) / 8;  // k-blocks of 64 incl. the bias chunk
    static_assert(NCH < KB * 8, "a spare operand chunk is needed for the bias columns");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w1_base = base;                                  // KB x 8 KB conv1 weights (resident)
    const uint32_t w2_base = w1_base + KB * SF_B_BYTES;             // 8 KB conv2_3x3_reduce weights (resident)
    const uint32_t p_tile = w2_base + SF_B_BYTES;                   // vertical-max tile / store staging (16 KB)
    const uint32_t o_tile = p_tile + A_STAGE_BYTES;                 // pooled tile = operand of the 1x1 GEMM (16 KB)
    const uint32_t stg = o_tile + A_STAGE_BYTES;                    // CIN x 11 x 264 fp16 staged input rows
    const uint32_t bar_base = stg + ((CIN * SF_RING * SF_SROW * 2 + 15) & ~15);
    const uint32_t w_bar = bar_base, mma_bar = bar_base + 8u, tmem_slot = bar_base + 16u;

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(w_bar, 1);
        mbar_init(mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW2)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    }
    // tensor memory: accumulator (64 columns) + conv1 operand tiles of KB x 32 columns (64 fp16 of K per k-block): two of
    // them (double buffered) when they fit, one for the 10-channel flow stem (K = 562: 288 columns)
    constexpr uint32_t TM_COLS = CIN == 1 ? 128u : (CIN == 3 ? 256u : 512u);
    constexpr uint32_t TM_A0 = 64u, TM_ASZ = (uint32_t)KB * 32u;
    constexpr bool DB = TM_A0 + 2u * TM_ASZ <= TM_COLS;  // operand double buffering
    static_assert(TM_A0 + TM_ASZ <= TM_COLS, "tensor memory plan");
    if (warp == 1) tmem_alloc(tmem_slot, TM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (tid == 0) {  // weights: resident for the life of the CTA
        mbar_expect_tx(w_bar, (KB + 1) * SF_B_BYTES);
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(w1_base + (uint32_t)kb * SF_B_BYTES, &tmW1, w_bar, kb * 64, 0);
        tma_load_2d(w2_base, &tmW2, w_bar, 0, 0);
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    // operand chunks beyond the last (c, r): chunk NCH = {1, 1, 0, ...} (multiplies the bias hi / lo weight columns), rest 0;
    // written once into both operand buffers (the MMAs only read them)
    for (int buf = 0; buf < (DB ? 2 : 1); ++buf)
        for (int ch = NCH; ch < KB * 8; ++ch)
            tmem_st4(taddr + TM_A0 + (uint32_t)buf * TM_ASZ + (uint32_t)ch * 4u, ch == NCH ? 0x3C003C00u : 0u, 0u, 0u, 0u);
    tmem_st_wait();

    const uint32_t idesc = umma_idesc_f16(SF_COUT);
    const uint32_t sw = (uint32_t)(tid & 7);
    uint32_t mma_phase = 0;
    bool weights_ready = false, store_pending = false;

    if (p.trace != nullptr && tid == 0 && blockIdx.x < 1024) {  // per-CTA start / end in %globaltimer ns (slots 128 ..)
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[128 + 2 * blockIdx.x] = t;
    }
    int tcount = -1;
    // Every CTA takes a CONTIGUOUS run of tiles, pooled row fastest: consecutive tiles are then vertically adjacent, and a
    // tile finds 7 of its 11 input rows already staged by its predecessor (staging was 3.7k of a tile's 10.7k cycles).
    const int t_base = p.tiles / (int)gridDim.x, t_rem = p.tiles % (int)gridDim.x;
    const int t_begin = (int)blockIdx.x * t_base + ((int)blockIdx.x < t_rem ? (int)blockIdx.x : t_rem);
    const int t_end = t_begin + t_base + ((int)blockIdx.x < t_rem ? 1 : 0);
    for (int tile = t_begin; tile < t_end; ++tile) {
        ++tcount;
        SF_STAMP(0);
        const int py = tile % p.ph;
        const int t2 = tile / p.ph;
        const int seg = t2 % p.segs, ni = t2 / p.segs;
        const int r_first = (tile > t_begin && py > 0) ? 7 : 0;  // rows 0..6 = rows 4..10 of the tile above (same column)
        const int slot0 = (4 * py + 9) % SF_RING;               // ring slot of the window's first input row 4*py - 3
        const int px0 = seg * SF_PW;
        const int ox0 = 2 * px0;
        int nq = p.oh - 2 * py;  // conv rows 2*py + q that exist
        nq = nq > 3 ? 3 : nq;

        // ---- 0. L2 prefetch of the rows the NEXT tile of this run will stage (its 4 new rows when it continues down the
        //         column).  In a real step the network input streams from HBM once: under ncu (cold) an RGB launch takes
        //         44.8 us against 38.0 us with the input L2-resident.
        if (tile + 1 < t_end && py + 1 < p.ph) {
            constexpr int LINES = (SF_SROW * (int)sizeof(TIN) + 127) / 128 + 1;
            const int c0 = 4 * seg * SF_PW - 4 < 0 ? 0 : 4 * seg * SF_PW - 4;
            const TIN* nx = reinterpret_cast<const TIN*>(p.x) + (size_t)ni * CIN * p.h * p.w;
            for (int idx = tid; idx < CIN * 4 * LINES; idx += 128) {
                const int r = idx / LINES, l = idx - r * LINES;
                const int c = r >> 2, iy = 4 * (py + 1) + 4 + (r & 3);
                const int col = c0 + l * (128 / (int)sizeof(TIN));
                if (iy < p.h && col < p.w) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + ((size_t)c * p.h + iy) * p.w + col));
            }
        }
        // ---- 1. stage the 11 x 264 input window of every channel as fp16 (coalesced, one column pair per thread) ----
        const TIN* xin = reinterpret_cast<const TIN*>(p.x) + (size_t)ni * CIN * p.h * p.w;
        const int ix_base = 2 * ox0 - 4;
        const int jx0 = ix_base + 2 * tid, jx1 = ix_base + 2 * (tid + 128);
        const bool c00 = (unsigned)jx0 < (unsigned)p.w, c01 = (unsigned)(jx0 + 1) < (unsigned)p.w;
        const bool c10 = tid < (SF_SROW / 2 - 128) && (unsigned)jx1 < (unsigned)p.w;
        const bool c11 = tid < (SF_SROW / 2 - 128) && (unsigned)(jx1 + 1) < (unsigned)p.w;
        if constexpr (sizeof(TIN) == 4) {
            // fp32 input.  The loads of CB channels are issued before the first conversion / store (the asm volatile stores
            // would otherwise fence the next channel's loads behind them).  Staging is bound by instruction issue (address
            // arithmetic + predication of ~130 loads per thread; tbn_debug_trace_stem: 3.7k of a tile's 10.7k cycles), so the
            // incremental case -- only the 4 new rows below the tile above -- is a separate, 4-row instantiation.
            constexpr int CB = CIN <= 3 ? CIN : 2;
            auto stage_rows = [&](auto r0_tag) {
                constexpr int R0 = decltype(r0_tag)::value;
                constexpr int NR = SF_ROWS - R0;
#pragma unroll 1
                for (int cb = 0; cb < CIN; cb += CB) {
                    float v0[CB][NR], v1[CB][NR], w0[CB][NR], w1[CB][NR];
#pragma unroll
                    for (int cc = 0; cc < CB; ++cc) {
                        const int c = cb + cc < CIN ? cb + cc : CIN - 1;
                        const int iy0 = 4 * py - 3 + R0;
                        const TIN* rowp = xin + ((long long)c * p.h + iy0) * p.w;  // (may point before the tensor: only
#pragma unroll                                                                     //  dereferenced for rows inside it)
                        for (int r = 0; r < NR; ++r) {
                            const bool row_in = (unsigned)(iy0 + r) < (unsigned)p.h;
                            const TIN* src = rowp + (long long)r * p.w;
                            v0[cc][r] = (row_in && c00) ? ld_in(src + jx0) : 0.f;
                            v1[cc][r] = (row_in && c01) ? ld_in(src + jx0 + 1) : 0.f;
                            w0[cc][r] = (row_in && c10) ? ld_in(src + jx1) : 0.f;
                            w1[cc][r] = (row_in && c11) ? ld_in(src + jx1 + 1) : 0.f;
                        }
                    }
#pragma unroll
                    for (int cc = 0; cc < CB; ++cc) {
                        const int c = cb + cc;
                        if (c < CIN) {
#pragma unroll
                            for (int r = 0; r < NR; ++r) {
                                int sl = slot0 + R0 + r;
                                sl -= sl >= SF_RING ? SF_RING : 0;
                                const uint32_t row = stg + (uint32_t)((c * SF_RING + sl) * SF_SROW) * 2u;
                                asm volatile("st.shared.u32 [%0], %1;" ::"r"(row + (uint32_t)tid * 4u), "r"(pack_h2(v0[cc][r], v1[cc][r]))
                                             : "memory");
                                if (tid < SF_SROW / 2 - 128)
                                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(row + (uint32_t)(tid + 128) * 4u),
                                                 "r"(pack_h2(w0[cc][r], w1[cc][r]))
                                                 : "memory");
                            }
                        }
                    }
                }
            };
            if (r_first == 0) stage_rows(std::integral_constant<int, 0>{});
            else stage_rows(std::integral_constant<int, 7>{});
        } else {
#pragma unroll(CIN <= 3 ? CIN : 2)
        for (int c = 0; c < CIN; ++c) {
            // packed fp16 pairs of this thread's two column pairs, all rows: loads first (all in flight), stores after
            uint32_t pa[SF_ROWS], pb[SF_ROWS];
            if constexpr (sizeof(TIN) == 4) {
                float v0[SF_ROWS], v1[SF_ROWS], w0[SF_ROWS], w1[SF_ROWS];
#pragma unroll
                for (int r = 0; r < SF_ROWS; ++r) {
                    const int iy = 4 * py - 3 + r;
                    const bool row_in = (unsigned)iy < (unsigned)p.h && r >= r_first;
                    const TIN* src = xin + ((size_t)c * p.h + (row_in ? iy : 0)) * p.w;
                    v0[r] = (row_in && c00) ? ld_in(src + jx0) : 0.f;
                    v1[r] = (row_in && c01) ? ld_in(src + jx0 + 1) : 0.f;
                    w0[r] = (row_in && c10) ? ld_in(src + jx1) : 0.f;
                    w1[r] = (row_in && c11) ? ld_in(src + jx1 + 1) : 0.f;
                }
#pragma unroll
                for (int r = 0; r < SF_ROWS; ++r) {
                    pa[r] = pack_h2(v0[r], v1[r]);
                    pb[r] = pack_h2(w0[r], w1[r]);
                }
            } else if ((p.w & 1) == 0) {
                // bf16 input, even row pitch: a column pair is one aligned 32-bit word
#pragma unroll
                for (int r = 0; r < SF_ROWS; ++r) {
                    const int iy = 4 * py - 3 + r;
                    const bool row_in = (unsigned)iy < (unsigned)p.h && r >= r_first;
                    const TIN* src = xin + ((size_t)c * p.h + (row_in ? iy : 0)) * p.w;
                    pa[r] = (row_in && c00) ? __ldg(reinterpret_cast<const uint32_t*>(src + jx0)) : 0u;
                    pb[r] = (row_in && c10) ? __ldg(reinterpret_cast<const uint32_t*>(src + jx1)) : 0u;
                }
#pragma unroll
                for (int r = 0; r < SF_ROWS; ++r) {
                    pa[r] = bf2_to_h2(pa[r]);
                    pb[r] = bf2_to_h2(pb[r]);
                }
            } else {
                unsigned short v0[SF_ROWS], v1[SF_ROWS], w0[SF_ROWS], w1[SF_ROWS];
#pragma unroll
                for (int r = 0; r < SF_ROWS; ++r) {
                    const int iy = 4 * py - 3 + r;
                    const bool row_in = (unsigned)iy < (unsigned)p.h && r >= r_first;
                    const unsigned short* src = reinterpret_cast<const unsigned short*>(xin + ((size_t)c * p.h + (row_in ? iy : 0)) * p.w);
                    v0[r] = (row_in && c00) ? __ldg(src + jx0) : (unsigned short)0;
                    v1[r] = (row_in && c01) ? __ldg(src + jx0 + 1) : (unsigned short)0;
                    w0[r] = (row_in && c10) ? __ldg(src + jx1) : (unsigned short)0;
                    w1[r] = (row_in && c11) ? __ldg(src + jx1 + 1) : (unsigned short)0;
                }
#pragma unroll
                for (int r = 0; r < SF_ROWS; ++r) {
                    pa[r] = bf2_to_h2((uint32_t)v0[r] | ((uint32_t)v1[r] << 16));
                    pb[r] = bf2_to_h2((uint32_t)w0[r] | ((uint32_t)w1[r] << 16));
                }
            }
#pragma unroll
            for (int r = 0; r < SF_ROWS; ++r) {
                if (r < r_first) continue;
                int sl = slot0 + r;
                sl -= sl >= SF_RING ? SF_RING : 0;
                const uint32_t row = stg + (uint32_t)((c * SF_RING + sl) * SF_SROW) * 2u;
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(row + (uint32_t)tid * 4u), "r"(pa[r]) : "memory");
                if (tid < SF_SROW / 2 - 128)
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(row + (uint32_t)(tid + 128) * 4u), "r"(pb[r]) : "memory");
            }
        }
        }
        if (tid == 0 && store_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging tile free
        __syncthreads();
        SF_STAMP(1);  // input window staged

        // ---- 2. three conv1 rows: operand tile -> MMA -> running maximum (starts at 0: ReLU + vertical max) ----------
        __half2 vmax[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) vmax[j] = __float2half2_rn(0.f);
        // operand row of conv row q -> TMEM buffer q & 1
        auto build = [&](int q) {
            const uint32_t dst = taddr + TM_A0 + (DB ? (uint32_t)(q & 1) * TM_ASZ : 0u);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int c = ch / 7, r = ch - c * 7;
                uint32_t v0, v1, v2, v3;
                int sl = slot0 + 2 * q + r;
                sl -= sl >= SF_RING ? SF_RING : 0;
                const uint32_t src = stg + (uint32_t)((c * SF_RING + sl) * SF_SROW) * 2u + (uint32_t)tid * 4u;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v0) : "r"(src));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v1) : "r"(src + 4u));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v2) : "r"(src + 8u));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v3) : "r"(src + 12u));
                tmem_st4(dst + (uint32_t)ch * 4u, v0, v1, v2, v3);
            }
            tmem_st_wait();
        };
        // conv1 MMAs of conv row q: K = NCH*8 + 2 (bias columns) in steps of 16 = 8 TMEM columns
        // (called by ALL lanes of warp 0, converged; the elected lane issues: a single divergent thread needs ~85 cycles per
        //  tcgen05.mma issue, measured with tbn_debug_trace_stem -- 900 cycles per conv row -- the converged warp far less)
        auto issue = [&](int q) {
            if (!weights_ready) {
                mbar_wait(w_bar, 0);
                weights_ready = true;
            }
            tc_fence_after();
            constexpr int NK16 = (NCH * 8 + 2 + 15) / 16;
            const uint32_t a_t = tmem_base + TM_A0 + (DB ? (uint32_t)(q & 1) * TM_ASZ : 0u);
            if (elect_one()) {
#pragma unroll
                for (int k16 = 0; k16 < NK16; ++k16) {
                    const uint64_t bdesc = umma_desc_sw128(w1_base + (uint32_t)(k16 >> 2) * SF_B_BYTES) + (uint64_t)(2 * (k16 & 3));
                    umma_bf16_ts(tmem_base, a_t + (uint32_t)k16 * 8u, bdesc, idesc, k16 != 0 ? 1u : 0u);
                }
                umma_commit(mma_bar);
            }
            __syncwarp();
        };
        build(0);
        tc_fence_before();
        __syncthreads();
        SF_STAMP(2);  // operand row 0 built
        if (warp == 0) issue(0);
        for (int q = 0; q < nq; ++q) {
            if (DB && q + 1 < nq) build(q + 1);  // overlaps the MMAs of row q
            SF_STAMP(3 + 3 * q);  // next operand row built
            mbar_wait(mma_bar, mma_phase);
            SF_STAMP(4 + 3 * q);  // MMAs of row q retired
            mma_phase ^= 1u;
            tc_fence_after();
            if (!DB && q + 1 < nq) build(q + 1);  // single operand buffer: free once the MMAs of row q have retired
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t acc[32];
                tmem_ld32(taddr + (uint32_t)(half * 32), acc);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    vmax[half * 16 + j] = __hmax2(
                        vmax[half * 16 + j], __floats2half2_rn(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1])));
            }
            tc_fence_before();
            __syncthreads();  // accumulator drained, operand row q+1 complete
            SF_STAMP(5 + 3 * q);  // accumulator of row q drained
            if (q + 1 < nq && warp == 0) issue(q + 1);
        }
        tc_fence_after();

        // ---- 3. vertical-max row of this conv column -> swizzled tile (columns outside the conv output count as 0) ----
        {
            const bool col_ok = ox0 + tid < p.ow;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = col_ok ? *reinterpret_cast<const uint32_t*>(&vmax[c8 * 4 + j]) : 0u;
                const uint32_t dst = p_tile + (uint32_t)tid * 128u + (uint32_t)(((uint32_t)c8 ^ sw) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                             : "memory");
            }
        }
        __syncthreads();
        // ---- 4. horizontal 3-tap maximum: pooled pixel j = conv columns 2j, 2j+1, 2j+2 -> operand tile of the 1x1 GEMM --
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int id = tid + k * 128;
            if (id < SF_PW * 8) {
                const int j = id >> 3, c8 = id & 7;
                uint4 m = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const uint32_t rowi = (uint32_t)(2 * j + t);
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                 : "r"(p_tile + rowi * 128u + ((((uint32_t)c8) ^ (rowi & 7u)) << 4)));
                    __half2* mm = reinterpret_cast<__half2*>(&m);
                    const __half2* vv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                    for (int e = 0; e < 4; ++e) mm[e] = __hmax2(mm[e], vv[e]);
                }
                const uint32_t dst = o_tile + (uint32_t)j * 128u + ((((uint32_t)c8) ^ ((uint32_t)j & 7u)) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(m.x), "r"(m.y), "r"(m.z), "r"(m.w) : "memory");
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        SF_STAMP(12);  // pooled tile ready
        // ---- 5. conv2_3x3_reduce: (63 pooled pixels x 64) @ W2^T on the tensor core --------------------------------------
        if (warp == 0) {
            if (!weights_ready) {  // (a tile with no conv rows cannot occur: ph rows always start inside the conv output)
                mbar_wait(w_bar, 0);
                weights_ready = true;
            }
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(o_tile);
            const uint64_t bdesc = umma_desc_sw128(w2_base);
            if (elect_one()) {
#pragma unroll
                for (int k16 = 0; k16 < 4; ++k16)
                    umma_bf16(tmem_base, adesc + (uint64_t)(2 * k16), bdesc + (uint64_t)(2 * k16), idesc, k16 != 0 ? 1u : 0u);
                umma_commit(mma_bar);
            }
            __syncwarp();
        }
        mbar_wait(mma_bar, mma_phase);
        mma_phase ^= 1u;
        tc_fence_after();
        SF_STAMP(13);  // 1x1 GEMM retired
        // ---- 6. + bias, ReLU, bf16 -> swizzled staging tile -> TMA store of the 63 x 64 result -------------------------------
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t acc[32];
            tmem_ld32(taddr + (uint32_t)(half * 32), acc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q8 = 0; q8 < 4; ++q8) {
                const int c8 = half * 4 + q8;
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias2 + c8 * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias2 + c8 * 8 + 4));
                const float bz[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float f0 = fmaxf(__uint_as_float(acc[q8 * 8 + 2 * j]) + bz[2 * j], 0.f);
                    const float f1 = fmaxf(__uint_as_float(acc[q8 * 8 + 2 * j + 1]) + bz[2 * j + 1], 0.f);
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
                    o[j] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                const uint32_t dst = p_tile + (uint32_t)tid * 128u + (uint32_t)(((uint32_t)c8 ^ sw) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                             : "memory");
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();  // staging complete; TMEM free for the next tile
        tc_fence_after();
        if (tid == 0) {
            tma_store_4d(&tmY, p_tile, 0, px0, py, ni);  // box {64 ch, 63 px, 1, 1}: clipped at the row end
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            store_pending = true;
        }
        SF_STAMP(14);  // epilogue done, store issued
    }

    if (p.trace != nullptr && tid == 0 && blockIdx.x < 1024) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[129 + 2 * blockIdx.x] = t;
    }
    if (warp == 0 && !weights_ready) mbar_wait(w_bar, 0);  // never exit with a TMA still in flight
    if (tid == 0) {
        if (store_pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TM_COLS);
}

unsigned long long* g_stem_trace = nullptr;

static int pool_out(int n) {  // MaxPool2d(3, 2, ceil_mode=True): the last window must start inside the input
    int o = (n - 3 + 1) / 2 + 1;
    if ((o - 1) * 2 >= n) --o;
    return o;
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_stem_fused_tc(const void* x, int x_dtype, const void* w1p, const void* w2p, const float* bias2, void* y, int n,
                                 int cin, int h, int w, int y_cs, int y_co, void* stream) {
    TBN_CHECK_ARG(x && w1p && w2p && bias2 && y && n > 0 && h >= 7 && w >= 7, "stem_fused_tc: bad arguments");
    TBN_CHECK_ARG(cin == 1 || cin == 3 || cin == 10, "stem_fused_tc: cin must be 1, 3 or 10 (got %d)", cin);
    TBN_CHECK_ARG(x_dtype == TBN_F32 || x_dtype == TBN_BF16, "stem_fused_tc: bad input dtype %d", x_dtype);
    TBN_CHECK_ARG(y_cs % 8 == 0 && y_co % 8 == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)w1p & 15) == 0 &&
                      ((uintptr_t)w2p & 15) == 0,
                  "stem_fused_tc: alignment");
    EncodeTiledFn encode = get_tensormap_encoder();
    TBN_CHECK_ARG(encode != nullptr, "stem_fused_tc: cuTensorMapEncodeTiled entry point not available");
    StemFusedParams p;
    p.x = x; p.bias2 = bias2;
    p.n = n; p.h = h; p.w = w;
    p.oh = (h + 6 - 7) / 2 + 1;
    p.ow = (w + 6 - 7) / 2 + 1;
    TBN_CHECK_ARG(p.oh >= 3 && p.ow >= 3, "stem_fused_tc: input too small");
    p.ph = pool_out(p.oh);
    p.pw = pool_out(p.ow);
    p.segs = (p.pw + SF_PW - 1) / SF_PW;
    const long long tiles = (long long)n * p.ph * p.segs;
    TBN_CHECK_ARG(tiles < (1LL << 31), "stem_fused_tc: too many tiles");
    p.tiles = (int)tiles;
    p.trace = g_stem_trace;
    const int nch = cin * 7, kb = (nch + 8) / 8;
    CUtensorMap tmW1, tmW2, tmY;
    {
        const long long kp = (long long)kb * 64;
        cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)SF_COUT};
        cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)SF_COUT};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&tmW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w1p), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TBN_CHECK_ARG(r == CUDA_SUCCESS, "stem_fused_tc: cuTensorMapEncodeTiled(W1) failed with %d", (int)r);
        cuuint64_t dims2[2] = {64, (cuuint64_t)SF_COUT};
        cuuint64_t strides2[1] = {128};
        r = encode(&tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w2p), dims2, strides2, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TBN_CHECK_ARG(r == CUDA_SUCCESS, "stem_fused_tc: cuTensorMapEncodeTiled(W2) failed with %d", (int)r);
    }
    {
        char* yb = reinterpret_cast<char*>(y) + (size_t)y_co * 2;
        const cuuint64_t pitch = (cuuint64_t)y_cs * 2;
        cuuint64_t dims[4] = {(cuuint64_t)SF_COUT, (cuuint64_t)p.pw, (cuuint64_t)p.ph, (cuuint64_t)n};
        cuuint64_t strides[3] = {pitch, pitch * p.pw, pitch * p.pw * p.ph};
        cuuint32_t box[4] = {64, SF_PW, 1, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, yb, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TBN_CHECK_ARG(r == CUDA_SUCCESS, "stem_fused_tc: cuTensorMapEncodeTiled(Y) failed with %d", (int)r);
    }
    // weights (conv1 k-blocks + conv2_3x3_reduce) + vertical-max / staging tile + pooled tile + staged input rows
    const size_t smem = (size_t)(kb + 1) * SF_B_BYTES + 2 * A_STAGE_BYTES + (size_t)cin * SF_RING * SF_SROW * 2 + 16 + 1024 + 64;
    int per_sm = (int)((228 * 1024) / (smem + 1024));  // 228 KB of shared memory per SM, 1 KB reserved per resident CTA
    const int max_per_sm = cin == 1 ? 4 : (cin == 3 ? 2 : 1);  // tensor-memory columns: 128 / 256 / 512 per CTA
    per_sm = per_sm < 1 ? 1 : (per_sm > max_per_sm ? max_per_sm : per_sm);
    long long grid = (long long)sm_count() * per_sm;
    if (grid > tiles) grid = tiles;
    typedef __nv_bfloat16 bf;
    static DeviceOnce once;
    int once_dev = 0;
    if (once.needed(&once_dev)) {
        const void* fns[6] = {(const void*)stem_fused_kernel<1, float>,  (const void*)stem_fused_kernel<1, bf>,
                              (const void*)stem_fused_kernel<3, float>,  (const void*)stem_fused_kernel<3, bf>,
                              (const void*)stem_fused_kernel<10, float>, (const void*)stem_fused_kernel<10, bf>};
        for (const void* fn : fns) {
            cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return cuda_fail(e, "stem_fused_tc smem opt-in");
        }
        once.done(once_dev);
    }
    const unsigned gr = (unsigned)grid;
    cudaStream_t st = as_stream(stream);
    const bool f32 = x_dtype == TBN_F32;
    if (cin == 1 && f32) stem_fused_kernel<1, float><<<gr, 128, smem, st>>>(tmW1, tmW2, tmY, p);
    else if (cin == 1) stem_fused_kernel<1, bf><<<gr, 128, smem, st>>>(tmW1, tmW2, tmY, p);
    else if (cin == 3 && f32) stem_fused_kernel<3, float><<<gr, 128, smem, st>>>(tmW1, tmW2, tmY, p);
    else if (cin == 3) stem_fused_kernel<3, bf><<<gr, 128, smem, st>>>(tmW1, tmW2, tmY, p);
    else if (f32) stem_fused_kernel<10, float><<<gr, 128, smem, st>>>(tmW1, tmW2, tmY, p);
    else stem_fused_kernel<10, bf><<<gr, 128, smem, st>>>(tmW1, tmW2, tmY, p);
    TBN_LAUNCH_CHECK("stem_fused_tc launch");
    return 0;
}

extern "C" int tbn_debug_trace_stem(void* dev_buf) {
    g_stem_trace = reinterpret_cast<unsigned long long*>(dev_buf);  // 8 tiles x 16 stamps + 1024 x (start, end) ns (null = off)
    return 0;
}
