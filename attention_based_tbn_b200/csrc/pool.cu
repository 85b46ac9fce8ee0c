// HBM-bound pooling kernels on NHWC activations: 8 channels (16 B bf16 / 32 B fp32) per thread,
// consecutive threads on consecutive channel vectors (fully coalesced), grid-stride loops sized
// to a multiple of the SM count.  Channel-slice in / channel-slice out (in-place concat).
#include <float.h>

#include "common.cuh"

namespace tbn {

struct PoolGeom {
    int n, h, w, c, x_cs, x_co, oh, ow, y_cs, y_co;
};

// 8-channel vectors: running max / sum helpers.  bf16 max is exact, so it stays packed (4 HMNMX2 per 16 B).
template <typename T>
struct Acc8;
template <>
struct Acc8<float> {
    float v[8];
    __device__ __forceinline__ void init(float x) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = x;
    }
    __device__ __forceinline__ void max_in(const float* p) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = fmaxf(v[0], a.x); v[1] = fmaxf(v[1], a.y); v[2] = fmaxf(v[2], a.z); v[3] = fmaxf(v[3], a.w);
        v[4] = fmaxf(v[4], b.x); v[5] = fmaxf(v[5], b.y); v[6] = fmaxf(v[6], b.z); v[7] = fmaxf(v[7], b.w);
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <>
struct Acc8<__nv_bfloat16> {
    __nv_bfloat162 v[4];
    __device__ __forceinline__ void init(float x) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __float2bfloat162_rn(x);
    }
    __device__ __forceinline__ void max_in(const __nv_bfloat16* p) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __hmax2(v[j], h[j]);
    }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(v); }
};

// MaxPool2d(3, stride 2, ceil_mode=True, padding 0): windows may overhang by one row/col.  Branch-free: overhanging
// taps are clamped onto the last row / column (a duplicate never changes a max), so the nine 16-byte loads of a
// thread are all in flight together; 32-bit index math (the host checks the element count).
template <typename T>
__global__ void maxpool3s2_kernel(const T* __restrict__ x, T* __restrict__ y, PoolGeom g) {
    const unsigned cv = (unsigned)g.c >> 3;
    const unsigned total = (unsigned)g.n * g.oh * g.ow * cv;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const unsigned c8 = (t % cv) * 8;
        unsigned p = t / cv;
        const unsigned ox = p % g.ow; p /= g.ow;
        const unsigned oy = p % g.oh;
        const unsigned ni = p / g.oh;
        const T* xin = x + (size_t)ni * g.h * g.w * g.x_cs + g.x_co + c8;
        Acc8<T> m;
        m.init(-FLT_MAX);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const unsigned iy = min(oy * 2 + r, (unsigned)g.h - 1);
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const unsigned ix = min(ox * 2 + s, (unsigned)g.w - 1);
                m.max_in(xin + (size_t)(iy * g.w + ix) * g.x_cs);
            }
        }
        m.store(y + ((size_t)(ni * g.oh + oy) * g.ow + ox) * g.y_cs + g.y_co + c8);
    }
}

// 3x3 stride 1 pad 1: average with count_include_pad (always /9) or max (-inf padding).  Branch-free: taps are
// clamped into the image; for the average a clamped (= padding) tap is multiplied by zero.
template <typename T, bool IS_MAX>
__global__ void pool3s1_kernel(const T* __restrict__ x, T* __restrict__ y, PoolGeom g) {
    const unsigned cv = (unsigned)g.c >> 3;
    const unsigned total = (unsigned)g.n * g.h * g.w * cv;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const unsigned c8 = (t % cv) * 8;
        unsigned p = t / cv;
        const int ox = (int)(p % g.w); p /= g.w;
        const int oy = (int)(p % g.h);
        const unsigned ni = p / g.h;
        const T* xin = x + (size_t)ni * g.h * g.w * g.x_cs + g.x_co + c8;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = IS_MAX ? -FLT_MAX : 0.f;
#pragma unroll
        for (int r = -1; r <= 1; ++r) {
            const int iy = min(max(oy + r, 0), g.h - 1);
#pragma unroll
            for (int s = -1; s <= 1; ++s) {
                const int ix = min(max(ox + s, 0), g.w - 1);
                const float keep = (IS_MAX || (iy == oy + r && ix == ox + s)) ? 1.f : 0.f;
                Vec8<T> v;
                v.load(xin + (size_t)(iy * g.w + ix) * g.x_cs);
                float f[8];
                v.get(f);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = IS_MAX ? fmaxf(a[j], f[j]) : fmaf(f[j], keep, a[j]);
            }
        }
        if (!IS_MAX) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] *= (1.0f / 9.0f);
        }
        Vec8<T> o;
        o.set(a);
        o.store(y + ((size_t)(ni * g.h + oy) * g.w + ox) * g.y_cs + g.y_co + c8);
    }
}

// avg 3x3/1 (count_include_pad) of a pre-activation, then + bias, ReLU (pool_proj commuted before the pool).
// A thread owns 8 channels of one column and walks RS output rows down it, carrying the horizontal 3-tap sums of the
// previous two input rows: 3 loads + 24 conversions per output instead of 9 + 72 (the kernel is instruction bound:
// the whole tensor sits in L2, it was just written by the block's fused GEMM).
template <typename T, int RS>
__global__ void avgpool3s1_bias_relu_kernel(const T* __restrict__ x, const float* __restrict__ bias, T* __restrict__ y,
                                            PoolGeom g) {
    const unsigned cv = (unsigned)g.c >> 3;
    const unsigned strips = ((unsigned)g.h + RS - 1) / RS;
    const unsigned total = (unsigned)g.n * strips * g.w * cv;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const unsigned c8 = (t % cv) * 8;
        unsigned p = t / cv;
        const int ox = (int)(p % g.w); p /= g.w;
        const int oy0 = (int)(p % strips) * RS;
        const unsigned ni = p / strips;
        const int oy1 = min(oy0 + RS, g.h);
        const T* xin = x + (size_t)ni * g.h * g.w * g.x_cs + g.x_co + c8;
        const int ixl = max(ox - 1, 0), ixr = min(ox + 1, g.w - 1);
        const float kl = ox > 0 ? 1.f : 0.f, kr = ox + 1 < g.w ? 1.f : 0.f;  // a clamped tap is zero padding
        auto hsum = [&](int iy, float* s) {
            if (iy < 0 || iy >= g.h) {
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] = 0.f;
                return;
            }
            const T* row = xin + (size_t)iy * g.w * g.x_cs;
            Vec8<T> vl, vc, vr;
            vl.load(row + (size_t)ixl * g.x_cs);
            vc.load(row + (size_t)ox * g.x_cs);
            vr.load(row + (size_t)ixr * g.x_cs);
            float fl[8], fc[8], fr[8];
            vl.get(fl); vc.get(fc); vr.get(fr);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = fmaf(fr[j], kr, fmaf(fl[j], kl, fc[j]));
        };
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 + 4));
        const float bz[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float ra[8], rb[8], rc[8];
        hsum(oy0 - 1, ra);
        hsum(oy0, rb);
        T* yout = y + ((size_t)(ni * g.h + oy0) * g.w + ox) * g.y_cs + g.y_co + c8;
        for (int oy = oy0; oy < oy1; ++oy, yout += (size_t)g.w * g.y_cs) {
            hsum(oy + 1, rc);
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a[j] = fmaxf(((ra[j] + rb[j]) + rc[j]) * (1.0f / 9.0f) + bz[j], 0.f);
                ra[j] = rb[j];
                rb[j] = rc[j];
            }
            Vec8<T> o;
            o.set(a);
            o.store(yout);
        }
    }
}

// split != nullptr: the pooled row is ALSO written as the bf16 [hi | hi | lo] operand of a head tensor-core GEMM (three
// blocks of c columns at split + row * split_ld: see tbn_split_bf16x3), so the head chain needs no separate split launch.
template <typename T, bool FREQ>
__global__ void avgpool_feat_kernel(const T* __restrict__ x, float* __restrict__ y, int n, int h, int w, int c,
                                    int y_ld, __nv_bfloat16* __restrict__ split, long long split_ld) {
    const int cv = c >> 3;
    const long long total = (long long)n * (FREQ ? w : 1) * cv;
    const float inv = FREQ ? 1.0f / h : 1.0f / ((float)h * w);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(t % cv) * 8;
        const long long q = t / cv;  // FREQ: n*w + x ; else n
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 0.f;
        if (FREQ) {
            const int ni = (int)(q / w), xi = (int)(q % w);
            for (int iy = 0; iy < h; ++iy) {
                Vec8<T> v;
                v.load(x + (((long long)ni * h + iy) * w + xi) * c + c8);
                float f[8];
                v.get(f);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] += f[j];
            }
        } else {
            const T* base = x + q * h * w * c + c8;
            for (int p = 0; p < h * w; ++p) {
                Vec8<T> v;
                v.load(base + (long long)p * c);
                float f[8];
                v.get(f);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] += f[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] *= inv;
        if (y != nullptr) {
            float* dst = y + q * y_ld + c8;
            *reinterpret_cast<float4*>(dst) = make_float4(a[0], a[1], a[2], a[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(a[4], a[5], a[6], a[7]);
        }
        if (split != nullptr) {
            __nv_bfloat16 hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                hi[j] = __float2bfloat16_rn(a[j]);
                lo[j] = __float2bfloat16_rn(a[j] - __bfloat162float(hi[j]));
            }
            __nv_bfloat16* o = split + q * split_ld + c8;
            const uint4 h4 = *reinterpret_cast<const uint4*>(hi), l4 = *reinterpret_cast<const uint4*>(lo);
            *reinterpret_cast<uint4*>(o) = h4;
            *reinterpret_cast<uint4*>(o + c) = h4;
            *reinterpret_cast<uint4*>(o + 2 * c) = l4;
        }
    }
}

// Global average over the h*w positions of a frame (the non-attended feature, bn_inception.py:32-35): one CTA per
// (frame, 128 channels), the positions split over 8 thread groups and combined through shared memory in a fixed order --
// the grid-stride kernel above gives this case only n*c/8 threads walking 49 positions each (27 us for 192 frames).
template <typename T>
__global__ void __launch_bounds__(128) avgpool_feat_global_kernel(const T* __restrict__ x, float* __restrict__ y, int hw, int c, int y_ld,
                                                                  __nv_bfloat16* __restrict__ split, long long split_ld) {
    __shared__ float red[8][16][8];
    const int tid = threadIdx.x, cg = tid & 15, part = tid >> 4;
    const int c8 = blockIdx.x * 128 + cg * 8;
    const long long q = blockIdx.y;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    if (c8 < c) {
        const T* base = x + q * hw * c + c8;
        for (int p = part; p < hw; p += 8) {
            Vec8<T> v;
            v.load(base + (long long)p * c);
            float f[8];
            v.get(f);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] += f[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[part][cg][j] = a[j];
    __syncthreads();
    if (part == 0 && c8 < c) {
        const float inv = 1.0f / (float)hw;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = red[0][cg][j];
#pragma unroll
            for (int k = 1; k < 8; ++k) t += red[k][cg][j];
            a[j] = t * inv;
        }
        if (y != nullptr) {
            float* dst = y + q * y_ld + c8;
            *reinterpret_cast<float4*>(dst) = make_float4(a[0], a[1], a[2], a[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(a[4], a[5], a[6], a[7]);
        }
        if (split != nullptr) {
            __nv_bfloat16 hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                hi[j] = __float2bfloat16_rn(a[j]);
                lo[j] = __float2bfloat16_rn(a[j] - __bfloat162float(hi[j]));
            }
            __nv_bfloat16* o = split + q * split_ld + c8;
            const uint4 h4 = *reinterpret_cast<const uint4*>(hi), l4 = *reinterpret_cast<const uint4*>(lo);
            *reinterpret_cast<uint4*>(o) = h4;
            *reinterpret_cast<uint4*>(o + c) = h4;
            *reinterpret_cast<uint4*>(o + 2 * c) = l4;
        }
    }
}

static int grid_for(long long total) {
    long long want = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

static int check_slices(const char* what, int c, int x_cs, int x_co, int y_cs, int y_co) {
    TBN_CHECK_ARG(c > 0 && c % 8 == 0 && x_cs % 8 == 0 && x_co % 8 == 0 && y_cs % 8 == 0 && y_co % 8 == 0,
                  "%s: channels, strides and offsets must be multiples of 8", what);
    TBN_CHECK_ARG(x_co + c <= x_cs && y_co + c <= y_cs, "%s: channel slice exceeds the channel stride", what);
    return 0;
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_maxpool3s2(const void* x, void* y, int n, int h, int w, int c, int x_cs, int x_co, int y_cs,
                              int y_co, int dtype, void* stream) {
    TBN_CHECK_ARG(x && y && n > 0 && h >= 3 && w >= 3, "maxpool3s2: bad arguments");
    if (int rc = check_slices("maxpool3s2", c, x_cs, x_co, y_cs, y_co)) return rc;
    PoolGeom g{n, h, w, c, x_cs, x_co, 0, 0, y_cs, y_co};
    g.oh = (h - 3 + 1) / 2 + 1;  // ceil((h-3)/2)+1
    g.ow = (w - 3 + 1) / 2 + 1;
    if ((g.oh - 1) * 2 >= h) --g.oh;  // last window must start inside the input
    if ((g.ow - 1) * 2 >= w) --g.ow;
    const long long total = (long long)n * g.oh * g.ow * (c / 8);
    TBN_CHECK_ARG(total < (1LL << 32) && (long long)n * h * w * x_cs < (1LL << 40), "maxpool3s2: tensor too large");
    if (dtype == TBN_F32)
        maxpool3s2_kernel<float><<<grid_for(total), 256, 0, as_stream(stream)>>>((const float*)x, (float*)y, g);
    else if (dtype == TBN_BF16)
        maxpool3s2_kernel<__nv_bfloat16>
            <<<grid_for(total), 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, g);
    else
        TBN_CHECK_ARG(false, "maxpool3s2: bad dtype %d", dtype);
    TBN_LAUNCH_CHECK("maxpool3s2 launch");
    return 0;
}

extern "C" int tbn_pool3s1(const void* x, void* y, int n, int h, int w, int c, int x_cs, int x_co, int y_cs,
                           int y_co, int mode, int dtype, void* stream) {
    TBN_CHECK_ARG(x && y && n > 0 && h > 0 && w > 0, "pool3s1: bad arguments");
    TBN_CHECK_ARG(mode == TBN_POOL_AVG || mode == TBN_POOL_MAX, "pool3s1: bad mode %d", mode);
    if (int rc = check_slices("pool3s1", c, x_cs, x_co, y_cs, y_co)) return rc;
    PoolGeom g{n, h, w, c, x_cs, x_co, h, w, y_cs, y_co};
    const long long total = (long long)n * h * w * (c / 8);
    TBN_CHECK_ARG(total < (1LL << 32), "pool3s1: tensor too large");
    const int blocks = grid_for(total);
    cudaStream_t st = as_stream(stream);
    if (dtype == TBN_F32) {
        if (mode == TBN_POOL_MAX)
            pool3s1_kernel<float, true><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, g);
        else
            pool3s1_kernel<float, false><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, g);
    } else if (dtype == TBN_BF16) {
        if (mode == TBN_POOL_MAX)
            pool3s1_kernel<__nv_bfloat16, true><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, g);
        else
            pool3s1_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, g);
    } else {
        TBN_CHECK_ARG(false, "pool3s1: bad dtype %d", dtype);
    }
    TBN_LAUNCH_CHECK("pool3s1 launch");
    return 0;
}

static int launch_avgpool_feat(const void* x, float* y, int n, int h, int w, int c, int y_ld, int freq_only, int dtype,
                               void* split, long long split_ld, void* stream);

extern "C" int tbn_avgpool_feat(const void* x, float* y, int n, int h, int w, int c, int y_ld, int freq_only,
                                int dtype, void* stream) {
    TBN_CHECK_ARG(y != nullptr, "avgpool_feat: null output");
    return launch_avgpool_feat(x, y, n, h, w, c, y_ld, freq_only, dtype, nullptr, 0, stream);
}

extern "C" int tbn_avgpool_feat_split(const void* x, float* y, void* split, long long split_ld, int n, int h, int w, int c,
                                      int y_ld, int freq_only, int dtype, void* stream) {
    TBN_CHECK_ARG(split != nullptr && split_ld >= 3LL * c && split_ld % 8 == 0 && ((uintptr_t)split & 15) == 0,
                  "avgpool_feat_split: bad split operand (needs 16-byte alignment, leading dimension >= 3c, %% 8)");
    return launch_avgpool_feat(x, y, n, h, w, c, y_ld, freq_only, dtype, split, split_ld, stream);
}

static int launch_avgpool_feat(const void* x, float* y, int n, int h, int w, int c, int y_ld, int freq_only, int dtype,
                               void* split, long long split_ld, void* stream) {
    TBN_CHECK_ARG(x && (y || split) && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && y_ld % 4 == 0 && y_ld >= c,
                  "avgpool_feat: bad arguments");
    __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(split);
    if (!freq_only && n <= 65535) {  // global pooling: CTA per (frame, 128 channels)
        const dim3 grid((unsigned)cdiv(c, 128), (unsigned)n);
        if (dtype == TBN_F32)
            avgpool_feat_global_kernel<float><<<grid, 128, 0, as_stream(stream)>>>((const float*)x, y, h * w, c, y_ld, sp, split_ld);
        else if (dtype == TBN_BF16)
            avgpool_feat_global_kernel<__nv_bfloat16><<<grid, 128, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, y, h * w, c, y_ld, sp,
                                                                                            split_ld);
        else
            TBN_CHECK_ARG(false, "avgpool_feat: bad dtype %d", dtype);
        TBN_LAUNCH_CHECK("avgpool_feat launch");
        return 0;
    }
    const long long total = (long long)n * (freq_only ? w : 1) * (c / 8);
    const int blocks = grid_for(total);
    cudaStream_t st = as_stream(stream);
    if (dtype == TBN_F32) {
        if (freq_only)
            avgpool_feat_kernel<float, true><<<blocks, 256, 0, st>>>((const float*)x, y, n, h, w, c, y_ld, sp, split_ld);
        else
            avgpool_feat_kernel<float, false><<<blocks, 256, 0, st>>>((const float*)x, y, n, h, w, c, y_ld, sp, split_ld);
    } else if (dtype == TBN_BF16) {
        if (freq_only)
            avgpool_feat_kernel<__nv_bfloat16, true><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, y, n, h, w, c, y_ld, sp, split_ld);
        else
            avgpool_feat_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, y, n, h, w, c, y_ld, sp, split_ld);
    } else {
        TBN_CHECK_ARG(false, "avgpool_feat: bad dtype %d", dtype);
    }
    TBN_LAUNCH_CHECK("avgpool_feat launch");
    return 0;
}

extern "C" int tbn_avgpool3s1_bias_relu(const void* x, const float* bias, void* y, int n, int h, int w, int c, int x_cs,
                                        int x_co, int y_cs, int y_co, int dtype, void* stream) {
    TBN_CHECK_ARG(x && y && bias && n > 0 && h > 0 && w > 0, "avgpool3s1_bias_relu: bad arguments");
    if (int rc = check_slices("avgpool3s1_bias_relu", c, x_cs, x_co, y_cs, y_co)) return rc;
    PoolGeom g{n, h, w, c, x_cs, x_co, h, w, y_cs, y_co};
    TBN_CHECK_ARG((long long)n * h * w * (c / 8) < (1LL << 32), "avgpool3s1_bias_relu: tensor too large");
    // output rows per thread: more rows = fewer loads per output ((RS + 2) * 3 / RS) but fewer threads; strips that
    // divide the map height leave no partly idle threads (28/14 -> 7, 32/16 -> 8; measured within 5 % of each other)
    const int rs = (h >= 14 && h % 7 == 0) ? 7 : ((h >= 16 && h % 8 == 0) ? 8 : 4);
    const long long total = (long long)n * ((h + rs - 1) / rs) * w * (c / 8);
#define TBN_AVGPOOL_LAUNCH(RS_)                                                                                              \
    do {                                                                                                                     \
        if (dtype == TBN_F32)                                                                                                \
            avgpool3s1_bias_relu_kernel<float, RS_>                                                                          \
                <<<grid_for(total), 256, 0, as_stream(stream)>>>((const float*)x, bias, (float*)y, g);                       \
        else if (dtype == TBN_BF16)                                                                                          \
            avgpool3s1_bias_relu_kernel<__nv_bfloat16, RS_>                                                                  \
                <<<grid_for(total), 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, bias, (__nv_bfloat16*)y, g);       \
        else                                                                                                                 \
            TBN_CHECK_ARG(false, "avgpool3s1_bias_relu: bad dtype %d", dtype);                                               \
    } while (0)
    if (rs == 7) TBN_AVGPOOL_LAUNCH(7);
    else if (rs == 8) TBN_AVGPOOL_LAUNCH(8);
    else TBN_AVGPOOL_LAUNCH(4);
#undef TBN_AVGPOOL_LAUNCH
    TBN_LAUNCH_CHECK("avgpool3s1_bias_relu launch");
    return 0;
}
