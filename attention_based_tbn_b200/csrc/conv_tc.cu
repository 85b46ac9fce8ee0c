// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation).
//
//   D[M = pixels, N = Cout] = A[M, K = taps x Cin] * W[K, N]   (+ bias, ReLU, channel-offset store)
//
// * A is never materialised: for every filter tap (r,s) and 64-channel slice the producer issues
//   ONE 4-D TMA box load {64 ch, BW, BH, BNI} of the NHWC activation at the shifted pixel
//   coordinate; out-of-bounds pixels (the conv padding) are zero-filled by the TMA unit.  The box
//   lands in shared memory as 128-byte rows with the 128B swizzle = the K-major SWIZZLE_128B
//   operand layout of tcgen05.mma, so no thread ever touches the operand bytes.
// * W is pre-packed [Cout][tap][Cin padded to 64] bf16 (BatchNorm scale folded in) and loaded by a
//   2-D TMA box {64, BN}.
// * One elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16) into a TMEM accumulator;
//   tcgen05.commit releases the smem stage / publishes the accumulator through mbarriers.
// * Epilogue: 4 warps read their 32 TMEM lanes (tcgen05.ld 32x32b), add the folded bias, ReLU,
//   convert and store straight into the channel slice of the block's concat buffer.
//
// Warp roles (128 threads): warp0 lane0 = TMA producer, warp1 lane0 = MMA issuer (+TMEM alloc),
// then all four warps run the epilogue.  Several CTAs are resident per SM so that the epilogue of
// one tile overlaps the main loop of another.
#include "tc_common.cuh"

#define TBN_MAX_SEGS 4

namespace tbn {

struct TcParams {
    int n, h, w;             // image batch / spatial size (output == input, stride 1)
    int bw, bh, bni;         // TMA pixel box
    int tiles_w, tiles_h;    // pixel tiles per image row / column (tiles over n implied)
    int n_tiles;             // tiles over Cout
    int cin;                 // input channels actually multiplied (multiple of 16)
    int cin_pad;             // cin rounded up to 64 (weight packing pitch per tap)
    int ksz, pad;            // filter size (1 or 3), padding
    int cout, bn_tile;
    int out_f32, bias_period;
    int num_kb, stages;
    unsigned tmem_cols;
    const float* bias;
    // output segments: columns [seg_end[s-1], seg_end[s]) go to seg_y[s] with channel stride seg_cs[s];
    // seg_co[s] already has the segment's first column subtracted (address = y + m*cs + co + col)
    int nseg;
    int seg_end[TBN_MAX_SEGS];
    int seg_cs[TBN_MAX_SEGS];
    int seg_co[TBN_MAX_SEGS];
    int seg_flags[TBN_MAX_SEGS];  // TBN_SEG_RELU | TBN_SEG_BIAS
    void* seg_y[TBN_MAX_SEGS];
};

__global__ void __launch_bounds__(128) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
    const uint32_t b_stage_bytes = (uint32_t)p.bn_tile * 128u;
    const uint32_t a_base = base;
    const uint32_t b_base = base + (uint32_t)p.stages * A_STAGE_BYTES;
    const uint32_t bar_base = b_base + (uint32_t)p.stages * b_stage_bytes;
    // barriers: full[s] at +0, empty[s] at +8*stages, accum at +16*stages, tmem slot after
    const uint32_t full0 = bar_base, empty0 = bar_base + 8u * p.stages, accum_bar = bar_base + 16u * p.stages;
    const uint32_t tmem_slot = accum_bar + 8u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int n_tile = blockIdx.x % p.n_tiles;
    int mt = blockIdx.x / p.n_tiles;
    const int tw = mt % p.tiles_w; mt /= p.tiles_w;
    const int th = mt % p.tiles_h;
    const int tn = mt / p.tiles_h;
    const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bni;
    const int col0 = n_tile * p.bn_tile;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int cchunks = p.cin_pad >> 6;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            const uint32_t stage_bytes =
                (uint32_t)(p.bw * p.bh * p.bni) * 128u + b_stage_bytes;
            int kb = 0;
            for (int tap = 0; tap < p.ksz * p.ksz; ++tap) {
                const int r = tap / p.ksz, s = tap - r * p.ksz;
                for (int cc = 0; cc < cchunks; ++cc, ++kb) {
                    const int st = kb % p.stages;
                    const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
                    mbar_wait(empty0 + 8u * st, ph ^ 1u);
                    mbar_expect_tx(full0 + 8u * st, stage_bytes);
                    tma_load_4d(a_base + (uint32_t)st * A_STAGE_BYTES, &tmA, full0 + 8u * st, cc * 64, w0 + s - p.pad,
                                h0 + r - p.pad, n0);
                    tma_load_2d(b_base + (uint32_t)st * b_stage_bytes, &tmB, full0 + 8u * st,
                                tap * p.cin_pad + cc * 64, col0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = umma_idesc_bf16(p.bn_tile);
            int kb = 0;
            for (int tap = 0; tap < p.ksz * p.ksz; ++tap) {
                for (int cc = 0; cc < cchunks; ++cc, ++kb) {
                    const int st = kb % p.stages;
                    const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
                    mbar_wait(full0 + 8u * st, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t adesc = umma_desc_sw128(a_base + (uint32_t)st * A_STAGE_BYTES);
                    const uint64_t bdesc = umma_desc_sw128(b_base + (uint32_t)st * b_stage_bytes);
                    int rem = p.cin - cc * 64;  // channels left in this slice (zero-filled tail is skipped)
                    const int nk16 = rem >= 64 ? 4 : (rem >> 4);
                    for (int k16 = 0; k16 < nk16; ++k16)
                        umma_bf16(tmem_base, adesc + (uint64_t)(2 * k16), bdesc + (uint64_t)(2 * k16), idesc,
                                  (kb | k16) != 0 ? 1u : 0u);
                    umma_commit(empty0 + 8u * st);  // frees the smem stage once these MMAs retire
                }
            }
            umma_commit(accum_bar);  // accumulator complete
        }
        __syncwarp();
    }

    // ===== epilogue: all four warps, warp q owns TMEM lanes 32q..32q+31 (= tile rows) =====
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const int row = warp * 32 + lane;
    const int wi = row % p.bw;
    const int hi = (row / p.bw) % p.bh;
    const int ni = row / (p.bw * p.bh);
    const bool row_ok = (ni < p.bni) && (w0 + wi < p.w) && (h0 + hi < p.h) && (n0 + ni < p.n);
    const long long mflat = ((long long)(n0 + ni) * p.h + (h0 + hi)) * p.w + (w0 + wi);
    const float* brow = p.bias + (p.bias_period > 1 ? (long long)(mflat % p.bias_period) * p.cout : 0);
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);

    int seg = 0;
    for (int c = 0; c < p.bn_tile; c += 16) {
        const int col = col0 + c;
        if (col >= p.cout) break;  // warp-uniform
        while (col >= p.seg_end[seg]) ++seg;  // 16-column chunks never straddle a segment
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row_ok) {
            float f[16];
            const int flags = p.seg_flags[seg];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
            if (flags & TBN_SEG_BIAS) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(brow + col + j));
                    f[j + 0] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                }
            }
            if (flags & TBN_SEG_RELU) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            const long long off = mflat * p.seg_cs[seg] + p.seg_co[seg] + col;
            if (p.out_f32) {
                float* dst = reinterpret_cast<float*>(p.seg_y[seg]) + off;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.seg_y[seg]) + off;
                uint4 o0, o1;
                __nv_bfloat162* h0p = reinterpret_cast<__nv_bfloat162*>(&o0);
                __nv_bfloat162* h1p = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    h0p[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                    h1p[j] = __floats2bfloat162_rn(f[8 + 2 * j], f[8 + 2 * j + 1]);
                }
                *reinterpret_cast<uint4*>(dst) = o0;
                *reinterpret_cast<uint4*>(dst + 8) = o1;
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols)
                     : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
EncodeTiledFn get_tensormap_encoder() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Pixel box {bw, bh, bni} with bw*bh*bni <= 128 that minimises the number of 128-row tiles.
static void pick_box(int n, int h, int w, int& bw, int& bh, int& bni) {
    long long best = -1;
    bw = bh = bni = 1;
    for (int cw = 1; cw <= w && cw <= 128; ++cw)
        for (int ch = 1; ch <= h && cw * ch <= 128; ++ch) {
            int cn = 128 / (cw * ch);
            if (cn > n) cn = n;
            const long long tiles = (long long)cdiv(w, cw) * cdiv(h, ch) * cdiv(n, cn);
            if (best < 0 || tiles < best || (tiles == best && cw * ch > bw * bh)) {
                best = tiles; bw = cw; bh = ch; bni = cn;
            }
        }
}

struct OutSeg {
    void* y;
    int y_cs, y_co, cols, flags;
};

static int launch_conv_tc(const void* x, const void* wp, const float* bias, const OutSeg* segs, int nseg, int n, int h,
                          int w, int cin, int x_cs, int x_co, int k, int out_dtype, int bias_period, void* stream) {
    TBN_CHECK_ARG(x && wp && bias && segs, "conv_tc: null pointer");
    TBN_CHECK_ARG(n > 0 && h > 0 && w > 0, "conv_tc: bad geometry");
    TBN_CHECK_ARG(k == 1 || k == 3, "conv_tc: k must be 1 or 3 (got %d); strided convs go through tbn_im2col", k);
    TBN_CHECK_ARG(cin % 16 == 0 && x_cs % 8 == 0 && x_co % 8 == 0, "conv_tc: cin %% 16, x_cs %% 8, x_co %% 8 required");
    TBN_CHECK_ARG(nseg >= 1 && nseg <= TBN_MAX_SEGS, "conv_tc: 1..%d output segments", TBN_MAX_SEGS);
    TBN_CHECK_ARG(out_dtype == TBN_BF16 || out_dtype == TBN_F32, "conv_tc: bad out_dtype");
    TBN_CHECK_ARG(bias_period >= 1, "conv_tc: bias_period >= 1");
    TBN_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 15) == 0, "conv_tc: pointers must be 16-byte aligned");
    EncodeTiledFn encode = get_tensormap_encoder();
    TBN_CHECK_ARG(encode != nullptr, "conv_tc: cuTensorMapEncodeTiled entry point not available");

    TcParams p;
    int cout = 0;
    for (int s = 0; s < nseg; ++s) {
        TBN_CHECK_ARG(segs[s].y && ((uintptr_t)segs[s].y & 15) == 0, "conv_tc: output pointers must be 16-byte aligned");
        TBN_CHECK_ARG(segs[s].cols > 0 && segs[s].cols % 16 == 0 && segs[s].y_cs % 8 == 0 && segs[s].y_co % 8 == 0,
                      "conv_tc: cout %% 16, y_cs %% 8, y_co %% 8 required");
        p.seg_y[s] = segs[s].y;
        p.seg_cs[s] = segs[s].y_cs;
        p.seg_co[s] = segs[s].y_co - cout;
        p.seg_flags[s] = segs[s].flags;
        cout += segs[s].cols;
        p.seg_end[s] = cout;
    }
    for (int s = nseg; s < TBN_MAX_SEGS; ++s) {
        p.seg_y[s] = nullptr; p.seg_cs[s] = 0; p.seg_co[s] = 0; p.seg_flags[s] = 0; p.seg_end[s] = 0x7fffffff;
    }
    p.nseg = nseg;
    // A 1x1 conv has no halo: flatten (n,h,w) into one pixel axis so that every tile is full.
    if (k == 1) {
        long long m = (long long)n * h * w;
        TBN_CHECK_ARG(m < (1LL << 31), "conv_tc: too many pixels");
        p.n = 1; p.h = 1; p.w = (int)m;
    } else {
        p.n = n; p.h = h; p.w = w;
    }
    pick_box(p.n, p.h, p.w, p.bw, p.bh, p.bni);
    p.tiles_w = cdiv(p.w, p.bw);
    p.tiles_h = cdiv(p.h, p.bh);
    const long long m_tiles = (long long)p.tiles_w * p.tiles_h * cdiv(p.n, p.bni);
    p.cin = cin;
    p.cin_pad = (cin + 63) / 64 * 64;
    p.ksz = k;
    p.pad = k / 2;
    p.cout = cout;
    // N tile: as wide as possible (A is then read once), narrowed while the grid underfills the GPU.
    int nt = cdiv(cout, 256);
    int bn = (cdiv(cout, nt) + 15) / 16 * 16;
    const int sms = sm_count();
    while (m_tiles * nt < sms && bn > 32) {  // bn strictly decreases: terminates
        bn -= 16;
        nt = cdiv(cout, bn);
        bn = (cdiv(cout, nt) + 15) / 16 * 16;  // rebalance the tiles for this tile count
        if (cdiv(cout, bn) != nt) bn -= 16;
    }
    nt = cdiv(cout, bn);
    p.bn_tile = bn;
    p.n_tiles = nt;
    p.out_f32 = (out_dtype == TBN_F32); p.bias_period = bias_period;
    p.num_kb = k * k * (p.cin_pad / 64);
    const int stage_bytes = A_STAGE_BYTES + bn * 128;
    int stages = (110 * 1024) / stage_bytes;
    if (stages < 2) stages = 2;
    if (stages > 6) stages = 6;
    if (stages > p.num_kb) stages = p.num_kb;
    p.stages = stages;
    p.tmem_cols = bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : 256;
    p.bias = bias;
    TBN_CHECK_ARG(m_tiles * nt < (1LL << 31), "conv_tc: grid too large");

    CUtensorMap tmA, tmB;
    {
        const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x) + x_co;
        cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.n};
        cuuint64_t strides[3] = {(cuuint64_t)x_cs * 2, (cuuint64_t)x_cs * 2 * p.w, (cuuint64_t)x_cs * 2 * p.w * p.h};
        cuuint32_t box[4] = {64, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bni};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)xb, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TBN_CHECK_ARG(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        const long long kp = (long long)k * k * p.cin_pad;
        cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)cout};
        cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)bn};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TBN_CHECK_ARG(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
    }

    const size_t smem = (size_t)stages * stage_bytes + 1024 + 16 * stages + 32;
    static thread_local bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "conv_tc smem opt-in");
        configured = true;
    }
    conv_tc_kernel<<<(unsigned)(m_tiles * nt), 128, smem, as_stream(stream)>>>(tmA, tmB, p);
    TBN_LAUNCH_CHECK("conv_tc launch");
    return 0;
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_conv_tc(const void* x, const void* wp, const float* bias, void* y, int n, int h, int w, int cin,
                           int x_cs, int x_co, int cout, int y_cs, int y_co, int k, int relu, int out_dtype,
                           int bias_period, void* stream) {
    OutSeg seg{y, y_cs, y_co, cout, TBN_SEG_BIAS | (relu ? TBN_SEG_RELU : 0)};
    return launch_conv_tc(x, wp, bias, &seg, 1, n, h, w, cin, x_cs, x_co, k, out_dtype, bias_period, stream);
}

extern "C" int tbn_conv_tc_multi(const void* x, const void* wp, const float* bias, const tbn_out_seg* segs_host,
                                 int nseg, int n, int h, int w, int cin, int x_cs, int x_co, int k, int out_dtype,
                                 void* stream) {
    TBN_CHECK_ARG(segs_host != nullptr && nseg >= 1 && nseg <= TBN_MAX_SEGS, "conv_tc_multi: bad segment list");
    OutSeg segs[TBN_MAX_SEGS];
    for (int s = 0; s < nseg; ++s)
        segs[s] = OutSeg{segs_host[s].y, segs_host[s].y_cs, segs_host[s].y_co, segs_host[s].cols, segs_host[s].flags};
    return launch_conv_tc(x, wp, bias, segs, nseg, n, h, w, cin, x_cs, x_co, k, out_dtype, 1, stream);
}
