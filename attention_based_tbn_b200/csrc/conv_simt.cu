// CUDA-core (fp32 FMA) implicit-GEMM conv + bias + ReLU, the input layout conversion and the
// patch gather for strided convs.  NHWC activations, channel-slice in / channel-slice out.
//
// conv_simt is the exact-precision path (fp32 parity mode, 1e-4 gate) and the on-device checker
// of the tcgen05 path; it is HBM/L2 + FMA bound and not the performance path.
#include "common.cuh"

namespace tbn {

struct ConvGeom {
    int n, h, w, cin, x_cs, x_co;
    int oh, ow, cout, y_cs, y_co;
    int k, stride, pad, relu;
    int K;          // k*k*cin
    long long M;    // n*oh*ow
};

constexpr int BM = 128, BN = 64, BK = 16;

template <typename T>
__device__ __forceinline__ void load4(const T* p, float* v);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float* v) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.x));
    float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float* v);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
    uint2 t;
    *reinterpret_cast<__nv_bfloat162*>(&t.x) = __floats2bfloat162_rn(v[0], v[1]);
    *reinterpret_cast<__nv_bfloat162*>(&t.y) = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
}

// One CTA: 128 output pixels x 64 output channels; 256 threads, 8x4 outputs each (two 4-pixel groups 64 apart); K in
// steps of 16 in ascending order, so every output is the same fmaf chain whatever the tile shape.
template <typename T, bool VEC4>
__global__ void __launch_bounds__(256) conv_simt_kernel(const T* __restrict__ x, const float* __restrict__ wk,
                                                        const float* __restrict__ bias, T* __restrict__ y,
                                                        ConvGeom g) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A-load role: this thread always gathers for pixel pm, the two k-quads kq*2 and kq*2+1 of a 16-wide k slab.
    const int pm = tid & 127, kq = tid >> 7;
    const long long m = m0 + pm;
    const bool m_ok = m < g.M;
    int iy0 = 0, ix0 = 0;
    const T* xin = x;
    if (m_ok) {
        const int ohw = g.oh * g.ow;
        const int ni = (int)(m / ohw);
        const int r = (int)(m - (long long)ni * ohw);
        const int oy = r / g.ow, ox = r - oy * g.ow;
        iy0 = oy * g.stride - g.pad;
        ix0 = ox * g.stride - g.pad;
        xin = x + (long long)ni * g.h * g.w * g.x_cs + g.x_co;
    }
    // B-load role
    const int brow = tid >> 4, bcol = (tid & 15) * 4;
    // compute role
    const int ty = tid >> 4, tx = tid & 15;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float av[2][4];
    float4 bv;
    // the next k slab is gathered into registers while the current one is multiplied (loads stay in flight over the FMAs)
    auto gather = [&](int k0) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int j = 0; j < 4; ++j) av[h2][j] = 0.f;
        // ---- gather A (im2col on the fly)
        if (m_ok) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int kk = k0 + (kq * 2 + h2) * 4;
                if (VEC4) {
                    if (kk < g.K) {
                        const int tap = kk / g.cin, c = kk - tap * g.cin;
                        const int r = tap / g.k, s = tap - r * g.k;
                        const int iy = iy0 + r, ix = ix0 + s;
                        if ((unsigned)iy < (unsigned)g.h && (unsigned)ix < (unsigned)g.w)
                            load4<T>(xin + ((long long)iy * g.w + ix) * g.x_cs + c, av[h2]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int kj = kk + j;
                        if (kj < g.K) {
                            const int tap = kj / g.cin, c = kj - tap * g.cin;
                            const int r = tap / g.k, s = tap - r * g.k;
                            const int iy = iy0 + r, ix = ix0 + s;
                            if ((unsigned)iy < (unsigned)g.h && (unsigned)ix < (unsigned)g.w)
                                av[h2][j] = to_f32<T>(xin[((long long)iy * g.w + ix) * g.x_cs + c]);
                        }
                    }
                }
            }
        }
        // ---- load B
        bv = make_float4(0.f, 0.f, 0.f, 0.f);
        {
            const int kr = k0 + brow, col = n0 + bcol;
            if (kr < g.K && col < g.cout) bv = *reinterpret_cast<const float4*>(wk + (long long)kr * g.cout + col);
        }
    };
    gather(0);
    for (int k0 = 0; k0 < g.K; k0 += BK) {
        __syncthreads();  // previous iteration's reads are done
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int j = 0; j < 4; ++j) As[(kq * 2 + h2) * 4 + j][pm] = av[h2][j];
        *reinterpret_cast<float4*>(&Bs[brow][bcol]) = bv;
        __syncthreads();
        if (k0 + BK < g.K) gather(k0 + BK);
#pragma unroll
        for (int q = 0; q < BK; ++q) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[q][64 + ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }

    const int col = n0 + tx * 4;
    if (col < g.cout) {
        float bz[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias) {
            const float4 t = *reinterpret_cast<const float4*>(bias + col);
            bz[0] = t.x; bz[1] = t.y; bz[2] = t.z; bz[3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long mo = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
            if (mo < g.M) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = acc[i][j] + bz[j];
                    if (g.relu) v[j] = fmaxf(v[j], 0.f);
                }
                store4<T>(y + mo * g.y_cs + g.y_co + col, v);
            }
        }
    }
}

// NCHW fp32 -> NHWC T, channels zero-padded to c_pad.  One thread per output pixel-channel
// vector; reads are coalesced along x for every channel plane, writes are contiguous.
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int n, int c, int h, int w,
                                    int c_pad) {
    const long long hw = (long long)h * w;
    const long long total = (long long)n * hw;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        const long long ni = p / hw, r = p - ni * hw;
        const float* src = x + ni * c * hw + r;
        T* dst = y + p * c_pad;
        for (int ci = 0; ci < c_pad; ++ci) dst[ci] = from_f32<T>(ci < c ? src[(long long)ci * hw] : 0.f);
    }
}

// Patch gather: A[m][kpad] (bf16), 8 consecutive columns per thread.  Column = tap*cin + c.
template <typename T, bool NCHW>
__global__ void im2col_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ a, ConvGeom g, int kpad) {
    const int kv = kpad >> 3;
    const long long total = g.M * kv;
    const int ohw = g.oh * g.ow;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long m = t / kv;
        const int k0 = (int)(t - m * kv) * 8;
        const int ni = (int)(m / ohw);
        const int r0 = (int)(m - (long long)ni * ohw);
        const int oy = r0 / g.ow, ox = r0 - oy * g.ow;
        const int iy0 = oy * g.stride - g.pad, ix0 = ox * g.stride - g.pad;
        float v[8];
        bool done = false;
        if (!NCHW && (g.cin & 7) == 0 && k0 < g.K) {
            // 8 consecutive channels of one tap: one vector load
            const int tap = k0 / g.cin, c = k0 - tap * g.cin;
            const int r = tap / g.k, s = tap - r * g.k;
            const int iy = iy0 + r, ix = ix0 + s;
            if ((unsigned)iy < (unsigned)g.h && (unsigned)ix < (unsigned)g.w) {
                Vec8<T> vv;
                vv.load(x + (((long long)ni * g.h + iy) * g.w + ix) * g.x_cs + g.x_co + c);
                vv.get(v);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.f;
            }
            done = true;
        }
        if (!done) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int kj = k0 + j;
                float val = 0.f;
                if (kj < g.K) {
                    const int tap = kj / g.cin, c = kj - tap * g.cin;
                    const int r = tap / g.k, s = tap - r * g.k;
                    const int iy = iy0 + r, ix = ix0 + s;
                    if ((unsigned)iy < (unsigned)g.h && (unsigned)ix < (unsigned)g.w) {
                        if (NCHW)
                            val = to_f32<T>(x[(((long long)ni * g.cin + c) * g.h + iy) * g.w + ix]);
                        else
                            val = to_f32<T>(x[(((long long)ni * g.h + iy) * g.w + ix) * g.x_cs + g.x_co + c]);
                    }
                }
                v[j] = val;
            }
        }
        Vec8<__nv_bfloat16> o;
        o.set(v);
        o.store(a + m * kpad + k0);
    }
}

static int make_geom(ConvGeom& g, int n, int h, int w, int cin, int x_cs, int x_co, int cout, int y_cs, int y_co,
                     int k, int stride, int pad, int relu) {
    TBN_CHECK_ARG(n > 0 && h > 0 && w > 0 && cin > 0 && k > 0 && stride > 0 && pad >= 0, "conv: bad geometry");
    g.n = n; g.h = h; g.w = w; g.cin = cin; g.x_cs = x_cs; g.x_co = x_co;
    g.oh = (h + 2 * pad - k) / stride + 1;
    g.ow = (w + 2 * pad - k) / stride + 1;
    TBN_CHECK_ARG(g.oh > 0 && g.ow > 0, "conv: empty output");
    g.cout = cout; g.y_cs = y_cs; g.y_co = y_co;
    g.k = k; g.stride = stride; g.pad = pad; g.relu = relu;
    g.K = k * k * cin;
    g.M = (long long)n * g.oh * g.ow;
    return 0;
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_conv_simt(const void* x, const float* wk, const float* bias, void* y, int n, int h, int w, int cin,
                             int x_cs, int x_co, int cout, int y_cs, int y_co, int k, int stride, int pad, int relu,
                             int dtype, void* stream) {
    ConvGeom g;
    if (int rc = make_geom(g, n, h, w, cin, x_cs, x_co, cout, y_cs, y_co, k, stride, pad, relu)) return rc;
    TBN_CHECK_ARG(x && wk && y, "conv_simt: null pointer");
    TBN_CHECK_ARG(cout % 4 == 0 && y_cs % 4 == 0 && y_co % 4 == 0, "conv_simt: cout/y_cs/y_co must be multiples of 4");
    TBN_CHECK_ARG(dtype == TBN_F32 || dtype == TBN_BF16, "conv_simt: bad dtype %d", dtype);
    const bool vec4 = (cin % 4 == 0) && (x_cs % 4 == 0) && (x_co % 4 == 0);
    dim3 grid(cdiv(g.M, BM), cdiv(cout, BN));
    cudaStream_t st = as_stream(stream);
    if (dtype == TBN_F32) {
        if (vec4)
            conv_simt_kernel<float, true><<<grid, 256, 0, st>>>((const float*)x, wk, bias, (float*)y, g);
        else
            conv_simt_kernel<float, false><<<grid, 256, 0, st>>>((const float*)x, wk, bias, (float*)y, g);
    } else {
        if (vec4)
            conv_simt_kernel<__nv_bfloat16, true>
                <<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, wk, bias, (__nv_bfloat16*)y, g);
        else
            conv_simt_kernel<__nv_bfloat16, false>
                <<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, wk, bias, (__nv_bfloat16*)y, g);
    }
    TBN_LAUNCH_CHECK("conv_simt launch");
    return 0;
}

extern "C" int tbn_nchw_to_nhwc(const float* x, void* y, int n, int c, int h, int w, int c_pad, int dtype,
                                void* stream) {
    TBN_CHECK_ARG(x && y && n > 0 && c > 0 && h > 0 && w > 0 && c_pad >= c, "nchw_to_nhwc: bad arguments");
    const long long total = (long long)n * h * w;
    const int blocks = (int)((total + 255) / 256 < (long long)sm_count() * 16 ? (total + 255) / 256 : (long long)sm_count() * 16);
    if (dtype == TBN_F32)
        nchw_to_nhwc_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(x, (float*)y, n, c, h, w, c_pad);
    else if (dtype == TBN_BF16)
        nchw_to_nhwc_kernel<__nv_bfloat16><<<blocks, 256, 0, as_stream(stream)>>>(x, (__nv_bfloat16*)y, n, c, h, w, c_pad);
    else
        TBN_CHECK_ARG(false, "nchw_to_nhwc: bad dtype %d", dtype);
    TBN_LAUNCH_CHECK("nchw_to_nhwc launch");
    return 0;
}

extern "C" int tbn_im2col(const void* x, void* a, int n, int h, int w, int cin, int x_cs, int x_co, int k, int stride,
                          int pad, int kpad, int x_dtype, int x_nchw, void* stream) {
    ConvGeom g;
    if (int rc = make_geom(g, n, h, w, cin, x_cs, x_co, 8, 8, 0, k, stride, pad, 0)) return rc;
    TBN_CHECK_ARG(x && a, "im2col: null pointer");
    TBN_CHECK_ARG(kpad % 8 == 0 && kpad >= g.K, "im2col: kpad must be a multiple of 8 and >= k*k*cin");
    TBN_CHECK_ARG(!x_nchw || x_dtype == TBN_F32, "im2col: NCHW input must be fp32");
    const long long total = g.M * (kpad / 8);
    long long want = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 32;
    const int blocks = (int)(want < cap ? want : cap);
    cudaStream_t st = as_stream(stream);
    __nv_bfloat16* ap = (__nv_bfloat16*)a;
    if (x_nchw)
        im2col_kernel<float, true><<<blocks, 256, 0, st>>>((const float*)x, ap, g, kpad);
    else if (x_dtype == TBN_F32)
        im2col_kernel<float, false><<<blocks, 256, 0, st>>>((const float*)x, ap, g, kpad);
    else if (x_dtype == TBN_BF16)
        im2col_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, ap, g, kpad);
    else
        TBN_CHECK_ARG(false, "im2col: bad dtype %d", x_dtype);
    TBN_LAUNCH_CHECK("im2col launch");
    return 0;
}
