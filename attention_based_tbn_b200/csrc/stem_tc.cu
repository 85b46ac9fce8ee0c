// Fused stem convolution (conv1_7x7_s2 + BatchNorm + ReLU) on tcgen05, reading the NCHW fp32
// network input directly (reference core/models/bn_inception_audio.py:35-39; model.py:211-213).
//
// The 7x7 / stride-2 conv has Cin = 1 (audio), 3 (RGB) or 10 (flow): too few channels for a TMA box,
// so the implicit-GEMM A tile is built by the CTA's threads instead ("software im2col"): thread t owns
// output pixel t of the 128-pixel tile (lanes = consecutive ox, so every gather instruction is a
// coalesced stride-2 read), converts to bf16 and stores 16-byte groups straight into the K-major
// SWIZZLE_128B operand layout (the XOR swizzle also makes those stores bank-conflict free).  The
// folded weights (64 x K, <= 64 KB) stay resident in shared memory for the whole persistent CTA; one
// thread issues the tcgen05.mma sequence into TMEM, then all threads run the bias + ReLU epilogue
// and write 64 bf16 channels (128 contiguous bytes) per pixel.  Nothing is materialised in HBM
// (the im2col + GEMM path this replaces wrote and re-read ~1.4 GB per 64-clip step).
#include "tc_common.cuh"

namespace tbn {

struct StemParams {
    const float* x;     // (n, cin, h, w) fp32
    const float* bias;  // (64)
    __nv_bfloat16* y;   // NHWC, channel stride y_cs, offset y_co
    int n, h, w, oh, ow;
    int y_cs, y_co;
    int kblocks;        // ceil(K / 64)
    int k16_total;      // ceil(K / 16)
    long long M;        // n*oh*ow
    int tiles;
};

constexpr int STEM_COUT = 64;
constexpr int STEM_B_BYTES = STEM_COUT * 128;  // one 64-wide k-block of the weights

template <int CIN>
__global__ void __launch_bounds__(128) stem_tc_kernel(const __grid_constant__ CUtensorMap tmW, const StemParams p) {
    constexpr int K = 49 * CIN;
    constexpr int KG = (K + 7) / 8;  // 16-byte groups per row actually gathered
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = base;                                            // kblocks x 8 KB
    const uint32_t a_base = base + (uint32_t)p.kblocks * STEM_B_BYTES;       // kblocks x 16 KB
    const uint32_t bar_base = a_base + (uint32_t)p.kblocks * A_STAGE_BYTES;
    const uint32_t w_bar = bar_base, mma_bar = bar_base + 8u, tmem_slot = bar_base + 16u;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(w_bar, 1);
        mbar_init(mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (tid == 0) {  // weights: resident for the life of the CTA
        mbar_expect_tx(w_bar, (uint32_t)p.kblocks * STEM_B_BYTES);
        for (int kb = 0; kb < p.kblocks; ++kb) tma_load_2d(b_base + (uint32_t)kb * STEM_B_BYTES, &tmW, w_bar, kb * 64, 0);
    }
    // zero the A columns beyond K once (k-groups KG .. kblocks*8-1 never change)
    for (int kg = KG; kg < p.kblocks * 8; ++kg) {
        const uint32_t addr = a_base + (uint32_t)(kg >> 3) * A_STAGE_BYTES + (uint32_t)tid * 128u +
                              (uint32_t)(((kg & 7) ^ (tid & 7)) << 4);
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
    }

    const uint32_t idesc = umma_idesc_bf16(STEM_COUT);
    const int ohw = p.oh * p.ow;
    const long long plane = (long long)p.h * p.w;
    uint32_t mma_phase = 0;
    bool weights_ready = false;

    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        // ---- 1. gather this thread's pixel row into the swizzled A tile --------------------------
        const long long m = (long long)tile * 128 + tid;
        const bool ok = m < p.M;
        int ni = 0, oy = 0, ox = 0;
        if (ok) {
            ni = (int)(m / ohw);
            const int r0 = (int)(m - (long long)ni * ohw);
            oy = r0 / p.ow;
            ox = r0 - oy * p.ow;
        }
        const int iy0 = oy * 2 - 3, ix0 = ox * 2 - 3;
        const float* xin = p.x + (long long)ni * CIN * plane;
        float v[8];
        int cnt = 0, kg = 0;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const int iy = iy0 + r;
            const bool row_in = ok && (unsigned)iy < (unsigned)p.h;
#pragma unroll
            for (int s = 0; s < 7; ++s) {
                const int ix = ix0 + s;
                const bool in = row_in && (unsigned)ix < (unsigned)p.w;
                const long long off = (long long)iy * p.w + ix;
#pragma unroll
                for (int c = 0; c < CIN; ++c) {
                    v[cnt] = in ? __ldg(xin + c * plane + off) : 0.f;
                    if (++cnt == 8) {
                        uint4 o;
                        __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                        for (int j = 0; j < 4; ++j) hp[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        const uint32_t addr = a_base + (uint32_t)(kg >> 3) * A_STAGE_BYTES + (uint32_t)tid * 128u +
                                              (uint32_t)(((kg & 7) ^ (tid & 7)) << 4);
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y),
                                     "r"(o.z), "r"(o.w)
                                     : "memory");
                        cnt = 0;
                        ++kg;
                    }
                }
            }
        }
        if (cnt != 0) {  // K % 8 != 0: flush the partial last group, zero padded
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j >= cnt) v[j] = 0.f;
            uint4 o;
            __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) hp[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            const uint32_t addr = a_base + (uint32_t)(kg >> 3) * A_STAGE_BYTES + (uint32_t)tid * 128u +
                                  (uint32_t)(((kg & 7) ^ (tid & 7)) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                         : "memory");
        }
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncthreads();

        // ---- 2. one thread issues the MMAs ---------------------------------------------------------
        if (tid == 0) {
            if (!weights_ready) {
                mbar_wait(w_bar, 0);
                weights_ready = true;
            }
            tc_fence_after();
            for (int kb = 0; kb < p.kblocks; ++kb) {
                const uint64_t adesc = umma_desc_sw128(a_base + (uint32_t)kb * A_STAGE_BYTES);
                const uint64_t bdesc = umma_desc_sw128(b_base + (uint32_t)kb * STEM_B_BYTES);
                const int rem = p.k16_total - kb * 4;
                const int nk16 = rem >= 4 ? 4 : rem;
                for (int k16 = 0; k16 < nk16; ++k16)
                    umma_bf16(tmem_base, adesc + (uint64_t)(2 * k16), bdesc + (uint64_t)(2 * k16), idesc,
                              (kb | k16) != 0 ? 1u : 0u);
            }
            umma_commit(mma_bar);
        }
        __syncwarp();

        // ---- 3. epilogue: TMEM lane = tile row = this thread's pixel ---------------------------------
        mbar_wait(mma_bar, mma_phase);
        mma_phase ^= 1u;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        __nv_bfloat16* dst = p.y + m * p.y_cs + p.y_co;
#pragma unroll
        for (int c = 0; c < STEM_COUT; c += 16) {
            uint32_t acc[16];
            tmem_ld16(taddr + (uint32_t)c, acc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (ok) {
                uint4 o0, o1;
                __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
                __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4 * j));
                    const float f0 = fmaxf(__uint_as_float(acc[4 * j + 0]) + b4.x, 0.f);
                    const float f1 = fmaxf(__uint_as_float(acc[4 * j + 1]) + b4.y, 0.f);
                    const float f2 = fmaxf(__uint_as_float(acc[4 * j + 2]) + b4.z, 0.f);
                    const float f3 = fmaxf(__uint_as_float(acc[4 * j + 3]) + b4.w, 0.f);
                    if (j < 2) {
                        h0[2 * j] = __floats2bfloat162_rn(f0, f1);
                        h0[2 * j + 1] = __floats2bfloat162_rn(f2, f3);
                    } else {
                        h1[2 * (j - 2)] = __floats2bfloat162_rn(f0, f1);
                        h1[2 * (j - 2) + 1] = __floats2bfloat162_rn(f2, f3);
                    }
                }
                *reinterpret_cast<uint4*>(dst + c) = o0;
                *reinterpret_cast<uint4*>(dst + c + 8) = o1;
            }
        }
        tc_fence_before();
        __syncthreads();  // TMEM and the A tile are free for the next tile
        tc_fence_after();
    }

    if (tid == 0 && !weights_ready) mbar_wait(w_bar, 0);  // never exit with a TMA still in flight
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 64);
}

// ---------------------------------------------------------------------------------------------------------
// Row-staged variant for cin = 1 (audio) and cin = 3 (RGB).  The per-pixel scalar gather above is latency bound
// (49*cin dependent global loads per thread, measured 80 % long-scoreboard stalls).  Here a tile is up to 128
// consecutive output pixels of ONE output row; the 7 input rows x (2*128 + 8) columns it touches are first copied,
// coalesced, into shared memory as bf16.  With the K layout
//        k = (c*7 + r)*8 + j,   j = 0..7  <->  input column 2*ox - 4 + j   (j = 0 carries a zero weight)
// the eight operand values of (pixel, c, r) are eight CONSECUTIVE staged bf16: four conflict-free 32-bit shared
// loads and one 16-byte store into the SWIZZLE_128B operand tile, no per-element conversion or packing.
// The epilogue writes bf16 rows into the (by then consumed) first operand k-block in the same swizzled layout and
// one thread issues a TMA tensor store of the 128 x 64 tile (clipped at the row end).
// ---------------------------------------------------------------------------------------------------------
constexpr int STEM_SROW = 264;  // staged bf16 per (channel, filter row): 2*128 + 8

template <int CIN>
__global__ void __launch_bounds__(128, CIN == 1 ? 6 : 2) stem_rows_kernel(const __grid_constant__ CUtensorMap tmW,
                                                        const __grid_constant__ CUtensorMap tmY, const StemParams p) {
    constexpr int NCH = CIN * 7;          // 16-byte operand chunks per pixel (one per (c, r))
    constexpr int KB = (NCH + 7) / 8;     // 64-wide k-blocks
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = base;                                   // KB x 8 KB weights (resident)
    const uint32_t a_base = base + KB * STEM_B_BYTES;               // KB x 16 KB operand tile
    const uint32_t stg = a_base + KB * A_STAGE_BYTES;               // NCH x STEM_SROW bf16 staged input rows
    const uint32_t bar_base = stg + ((NCH * STEM_SROW * 2 + 15) & ~15);
    const uint32_t w_bar = bar_base, mma_bar = bar_base + 8u, tmem_slot = bar_base + 16u;

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(w_bar, 1);
        mbar_init(mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (tid == 0) {  // weights: resident for the life of the CTA
        mbar_expect_tx(w_bar, KB * STEM_B_BYTES);
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(b_base + (uint32_t)kb * STEM_B_BYTES, &tmW, w_bar, kb * 64, 0);
    }
    // operand chunks beyond the last (c, r) never change: chunk NCH holds {1, 1, 0, ...} (it multiplies the weight
    // columns that carry the folded bias as bf16 hi + lo: the bias add rides on the MMA), the rest are zero
    static_assert(NCH < KB * 8, "a spare operand chunk is needed for the bias columns");
    for (int ch = NCH; ch < KB * 8; ++ch) {
        const uint32_t addr = a_base + (uint32_t)(ch >> 3) * A_STAGE_BYTES + (uint32_t)tid * 128u +
                              (uint32_t)(((ch & 7) ^ (tid & 7)) << 4);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %2, %2};" ::"r"(addr), "r"(ch == NCH ? 0x3F803F80u : 0u), "r"(0u)
                     : "memory");
    }

    const uint32_t idesc = umma_idesc_bf16(STEM_COUT);
    const int segs = (p.ow + 127) >> 7;
    uint32_t mma_phase = 0;
    bool weights_ready = false, store_pending = false;

    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int seg = tile % segs;
        const int t2 = tile / segs;
        const int oy = t2 % p.oh, ni = t2 / p.oh;
        const int ox0 = seg << 7;
        // ---- 1. stage the 7 x 264 input window of every channel as bf16 (coalesced, two columns per thread) ----
        const float* xin = p.x + (size_t)ni * CIN * p.h * p.w;
        const int ix_base = 2 * ox0 - 4;
        // thread t stages column pair t of every (c, r) row (threads 0..3 also pair 128 + t): the column bounds are the
        // same for all rows, the row pointer / row bound is uniform -> ~8 instructions per staged pair
        const int jx0 = ix_base + 2 * tid, jx1 = ix_base + 2 * (tid + 128);
        const bool c00 = (unsigned)jx0 < (unsigned)p.w, c01 = (unsigned)(jx0 + 1) < (unsigned)p.w;
        const bool c10 = tid < (STEM_SROW / 2 - 128) && (unsigned)jx1 < (unsigned)p.w;
        const bool c11 = tid < (STEM_SROW / 2 - 128) && (unsigned)(jx1 + 1) < (unsigned)p.w;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
            float v0[7], v1[7], w0[7], w1[7];
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                const int iy = 2 * oy - 3 + r;
                const bool row_in = (unsigned)iy < (unsigned)p.h;
                const float* src = xin + ((size_t)c * p.h + (row_in ? iy : 0)) * p.w;
                v0[r] = (row_in && c00) ? __ldg(src + jx0) : 0.f;
                v1[r] = (row_in && c01) ? __ldg(src + jx0 + 1) : 0.f;
                w0[r] = (row_in && c10) ? __ldg(src + jx1) : 0.f;
                w1[r] = (row_in && c11) ? __ldg(src + jx1 + 1) : 0.f;
            }
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                const uint32_t row = stg + (uint32_t)((c * 7 + r) * STEM_SROW) * 2u;
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0[r], v1[r]);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(row + (uint32_t)tid * 4u), "r"(*reinterpret_cast<const uint32_t*>(&h2))
                             : "memory");
                if (tid < STEM_SROW / 2 - 128) {
                    const __nv_bfloat162 g2 = __floats2bfloat162_rn(w0[r], w1[r]);
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(row + (uint32_t)(tid + 128) * 4u),
                                 "r"(*reinterpret_cast<const uint32_t*>(&g2))
                                 : "memory");
                }
            }
        }
        if (tid == 0 && store_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // A tile is free again
        __syncthreads();
        // ---- 2. operand rows: pixel tid, chunk (c, r) = staged bf16 [2*tid, 2*tid + 8) ---------------------------
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            uint32_t v0, v1, v2, v3;
            const uint32_t src = stg + (uint32_t)(ch * STEM_SROW) * 2u + (uint32_t)tid * 4u;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v0) : "r"(src));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v1) : "r"(src + 4u));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v2) : "r"(src + 8u));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v3) : "r"(src + 12u));
            const uint32_t dst = a_base + (uint32_t)(ch >> 3) * A_STAGE_BYTES + (uint32_t)tid * 128u +
                                 (uint32_t)(((ch & 7) ^ (tid & 7)) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
        }
        if (NCH < 8) {  // the bias chunk lives in k-block 0, which the previous tile's epilogue used as its staging tile
            const uint32_t dst = a_base + (uint32_t)tid * 128u + (uint32_t)(((NCH & 7) ^ (tid & 7)) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %2, %2};" ::"r"(dst), "r"(0x3F803F80u), "r"(0u) : "memory");
        }
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncthreads();
        // ---- 3. one thread issues the MMAs ------------------------------------------------------------------------
        if (tid == 0) {
            if (!weights_ready) {
                mbar_wait(w_bar, 0);
                weights_ready = true;
            }
            tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                const uint64_t adesc = umma_desc_sw128(a_base + (uint32_t)kb * A_STAGE_BYTES);
                const uint64_t bdesc = umma_desc_sw128(b_base + (uint32_t)kb * STEM_B_BYTES);
                const int rem16 = (NCH * 8 + 2 - kb * 64 + 15) / 16;  // + the two bias columns
                const int nk16 = rem16 >= 4 ? 4 : rem16;
                for (int k16 = 0; k16 < nk16; ++k16)
                    umma_bf16(tmem_base, adesc + (uint64_t)(2 * k16), bdesc + (uint64_t)(2 * k16), idesc,
                              (kb | k16) != 0 ? 1u : 0u);
            }
            umma_commit(mma_bar);
        }
        __syncwarp();
        // ---- 4. epilogue: TMEM lane = this thread's pixel; bf16 row -> swizzled staging (operand k-block 0) -> TMA store
        mbar_wait(mma_bar, mma_phase);
        mma_phase ^= 1u;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t acc[32];
            tmem_ld32(taddr + (uint32_t)(half * 32), acc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const __nv_bfloat162 zero2 = __float2bfloat162_rn(0.f);
#pragma unroll
            for (int q8 = 0; q8 < 4; ++q8) {
                const int c8 = half * 4 + q8;
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // the accumulator already contains the bias: convert, ReLU on packed bf16
                    const __nv_bfloat162 h2 = __hmax2(
                        __floats2bfloat162_rn(__uint_as_float(acc[q8 * 8 + 2 * j]), __uint_as_float(acc[q8 * 8 + 2 * j + 1])), zero2);
                    o[j] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                const uint32_t dst = a_base + (uint32_t)tid * 128u + (uint32_t)((c8 ^ (tid & 7)) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                             : "memory");
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();  // staging complete; TMEM free for the next tile
        tc_fence_after();
        if (tid == 0) {
            tma_store_4d(&tmY, a_base, 0, ox0, oy, ni);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            store_pending = true;
        }
    }

    if (tid == 0) {
        if (!weights_ready) mbar_wait(w_bar, 0);  // never exit with a TMA still in flight
        if (store_pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 64);
}

static int launch_stem_rows(const float* x, const void* wp, const float* bias, void* y, int n, int cin, int h, int w,
                            int y_cs, int y_co, void* stream, EncodeTiledFn encode) {
    StemParams p;
    p.x = x; p.bias = bias; p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.n = n; p.h = h; p.w = w;
    p.oh = (h + 6 - 7) / 2 + 1;
    p.ow = (w + 6 - 7) / 2 + 1;
    p.y_cs = y_cs; p.y_co = y_co;
    const int nch = cin * 7;
    p.kblocks = (nch + 7) / 8;
    p.k16_total = 0;
    p.M = (long long)n * p.oh * p.ow;
    const long long tiles = (long long)n * p.oh * ((p.ow + 127) / 128);
    TBN_CHECK_ARG(tiles < (1LL << 31), "stem_tc: too many tiles");
    p.tiles = (int)tiles;
    CUtensorMap tmW, tmY;
    {
        const long long kp = (long long)p.kblocks * 64;
        cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)STEM_COUT};
        cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)STEM_COUT};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TBN_CHECK_ARG(r == CUDA_SUCCESS, "stem_tc: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    {
        char* yb = reinterpret_cast<char*>(y) + (size_t)y_co * 2;
        const cuuint64_t pitch = (cuuint64_t)y_cs * 2;
        cuuint64_t dims[4] = {(cuuint64_t)STEM_COUT, (cuuint64_t)p.ow, (cuuint64_t)p.oh, (cuuint64_t)n};
        cuuint64_t strides[3] = {pitch, pitch * p.ow, pitch * p.ow * p.oh};
        cuuint32_t box[4] = {64, 128, 1, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, yb, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TBN_CHECK_ARG(r == CUDA_SUCCESS, "stem_tc: cuTensorMapEncodeTiled(Y) failed with %d", (int)r);
    }
    const size_t smem = (size_t)p.kblocks * (STEM_B_BYTES + A_STAGE_BYTES) + (size_t)nch * STEM_SROW * 2 + 16 + 1024 + 64;
    int per_sm = (int)((220 * 1024) / smem);
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    long long grid = (long long)sm_count() * per_sm;
    if (grid > tiles) grid = tiles;
    static DeviceOnce once;
    int once_dev = 0;
    if (once.needed(&once_dev)) {
        cudaError_t e = cudaFuncSetAttribute(stem_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_rows_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "stem_tc smem opt-in");
        once.done(once_dev);
    }
    if (cin == 1)
        stem_rows_kernel<1><<<(unsigned)grid, 128, smem, as_stream(stream)>>>(tmW, tmY, p);
    else
        stem_rows_kernel<3><<<(unsigned)grid, 128, smem, as_stream(stream)>>>(tmW, tmY, p);
    TBN_LAUNCH_CHECK("stem_tc (rows) launch");
    return 0;
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_stem_tc(const float* x, const void* wp, const float* bias, void* y, int n, int cin, int h, int w,
                           int y_cs, int y_co, void* stream) {
    TBN_CHECK_ARG(x && wp && bias && y && n > 0 && h > 0 && w > 0, "stem_tc: bad arguments");
    TBN_CHECK_ARG(cin == 1 || cin == 3 || cin == 10, "stem_tc: cin must be 1, 3 or 10 (got %d)", cin);
    TBN_CHECK_ARG(y_cs % 8 == 0 && y_co % 8 == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)wp & 15) == 0,
                  "stem_tc: alignment");
    EncodeTiledFn encode = get_tensormap_encoder();
    TBN_CHECK_ARG(encode != nullptr, "stem_tc: cuTensorMapEncodeTiled entry point not available");
    if (cin == 1 || cin == 3) return launch_stem_rows(x, wp, bias, y, n, cin, h, w, y_cs, y_co, stream, encode);
    StemParams p;
    p.x = x; p.bias = bias; p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.n = n; p.h = h; p.w = w;
    p.oh = (h + 6 - 7) / 2 + 1;
    p.ow = (w + 6 - 7) / 2 + 1;
    p.y_cs = y_cs; p.y_co = y_co;
    const int K = 49 * cin;
    p.kblocks = (K + 63) / 64;
    p.k16_total = (K + 15) / 16;
    p.M = (long long)n * p.oh * p.ow;
    const long long tiles = (p.M + 127) / 128;
    TBN_CHECK_ARG(tiles < (1LL << 31), "stem_tc: too many tiles");
    p.tiles = (int)tiles;

    CUtensorMap tmW;
    {
        const long long kp = (long long)p.kblocks * 64;  // packed pitch: K rounded up to 64
        cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)STEM_COUT};
        cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)STEM_COUT};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TBN_CHECK_ARG(r == CUDA_SUCCESS, "stem_tc: cuTensorMapEncodeTiled failed with %d", (int)r);
    }
    const size_t smem = (size_t)p.kblocks * (STEM_B_BYTES + A_STAGE_BYTES) + 1024 + 64;
    // the gather is latency bound (49*cin scalar loads per pixel): as many CTAs per SM as shared memory and the
    // 512 TMEM columns (64 per CTA) allow
    int per_sm = (int)((220 * 1024) / smem);
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    long long grid = (long long)sm_count() * per_sm;
    if (grid > tiles) grid = tiles;
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaSuccess;
    static DeviceOnce once;
    int once_dev = 0;
    if (once.needed(&once_dev)) {
        e = cudaFuncSetAttribute(stem_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_tc_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "stem_tc smem opt-in");
        once.done(once_dev);
    }
    if (cin == 1)
        stem_tc_kernel<1><<<(unsigned)grid, 128, smem, st>>>(tmW, p);
    else if (cin == 3)
        stem_tc_kernel<3><<<(unsigned)grid, 128, smem, st>>>(tmW, p);
    else
        stem_tc_kernel<10><<<(unsigned)grid, 128, smem, st>>>(tmW, p);
    TBN_LAUNCH_CHECK("stem_tc launch");
    return 0;
}
