// HBM-bound pooling kernels on NHWC activations: 8 channels (16 B bf16 / 32 B fp32) per thread,
// consecutive threads on consecutive channel vectors (fully coalesced), grid-stride loops sized
// to a multiple of the SM count.  Channel-slice in / channel-slice out (in-place concat).
#include <float.h>

#include "common.cuh"

namespace tbn {

struct PoolGeom {
    int n, h, w, c, x_cs, x_co, oh, ow, y_cs, y_co;
};

// MaxPool2d(3, stride 2, ceil_mode=True, padding 0): windows may overhang by one row/col.
template <typename T>
__global__ void maxpool3s2_kernel(const T* __restrict__ x, T* __restrict__ y, PoolGeom g) {
    const int cv = g.c >> 3;
    const long long total = (long long)g.n * g.oh * g.ow * cv;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(t % cv) * 8;
        long long p = t / cv;
        const int ox = (int)(p % g.ow); p /= g.ow;
        const int oy = (int)(p % g.oh);
        const int ni = (int)(p / g.oh);
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -FLT_MAX;
        const int iy0 = oy * 2, ix0 = ox * 2;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int iy = iy0 + r;
            if (iy >= g.h) break;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int ix = ix0 + s;
                if (ix >= g.w) break;
                Vec8<T> v;
                v.load(x + (((long long)ni * g.h + iy) * g.w + ix) * g.x_cs + g.x_co + c8);
                float f[8];
                v.get(f);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
            }
        }
        Vec8<T> o;
        o.set(m);
        o.store(y + (((long long)ni * g.oh + oy) * g.ow + ox) * g.y_cs + g.y_co + c8);
    }
}

// 3x3 stride 1 pad 1: average with count_include_pad (always /9) or max (-inf padding).
template <typename T, bool IS_MAX>
__global__ void pool3s1_kernel(const T* __restrict__ x, T* __restrict__ y, PoolGeom g) {
    const int cv = g.c >> 3;
    const long long total = (long long)g.n * g.h * g.w * cv;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(t % cv) * 8;
        long long p = t / cv;
        const int ox = (int)(p % g.w); p /= g.w;
        const int oy = (int)(p % g.h);
        const int ni = (int)(p / g.h);
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = IS_MAX ? -FLT_MAX : 0.f;
#pragma unroll
        for (int r = -1; r <= 1; ++r) {
            const int iy = oy + r;
            if ((unsigned)iy >= (unsigned)g.h) continue;
#pragma unroll
            for (int s = -1; s <= 1; ++s) {
                const int ix = ox + s;
                if ((unsigned)ix >= (unsigned)g.w) continue;
                Vec8<T> v;
                v.load(x + (((long long)ni * g.h + iy) * g.w + ix) * g.x_cs + g.x_co + c8);
                float f[8];
                v.get(f);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = IS_MAX ? fmaxf(a[j], f[j]) : a[j] + f[j];
            }
        }
        if (!IS_MAX) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] *= (1.0f / 9.0f);
        }
        Vec8<T> o;
        o.set(a);
        o.store(y + (((long long)ni * g.h + oy) * g.w + ox) * g.y_cs + g.y_co + c8);
    }
}

// avg 3x3/1 (count_include_pad) of a pre-activation, then + bias, ReLU (pool_proj commuted before the pool).
template <typename T>
__global__ void avgpool3s1_bias_relu_kernel(const T* __restrict__ x, const float* __restrict__ bias, T* __restrict__ y,
                                            PoolGeom g) {
    const int cv = g.c >> 3;
    const long long total = (long long)g.n * g.h * g.w * cv;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(t % cv) * 8;
        long long p = t / cv;
        const int ox = (int)(p % g.w); p /= g.w;
        const int oy = (int)(p % g.h);
        const int ni = (int)(p / g.h);
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
        for (int r = -1; r <= 1; ++r) {
            const int iy = oy + r;
            if ((unsigned)iy >= (unsigned)g.h) continue;
#pragma unroll
            for (int s = -1; s <= 1; ++s) {
                const int ix = ox + s;
                if ((unsigned)ix >= (unsigned)g.w) continue;
                Vec8<T> v;
                v.load(x + (((long long)ni * g.h + iy) * g.w + ix) * g.x_cs + g.x_co + c8);
                float f[8];
                v.get(f);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] += f[j];
            }
        }
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 + 4));
        const float bz[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j] * (1.0f / 9.0f) + bz[j], 0.f);
        Vec8<T> o;
        o.set(a);
        o.store(y + (((long long)ni * g.h + oy) * g.w + ox) * g.y_cs + g.y_co + c8);
    }
}

// Feature pooling -> fp32.  FREQ: mean over H for every (n, x); else mean over (H, W).
template <typename T, bool FREQ>
__global__ void avgpool_feat_kernel(const T* __restrict__ x, float* __restrict__ y, int n, int h, int w, int c,
                                    int y_ld) {
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
        float* dst = y + q * y_ld + c8;
        *reinterpret_cast<float4*>(dst) = make_float4(a[0] * inv, a[1] * inv, a[2] * inv, a[3] * inv);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(a[4] * inv, a[5] * inv, a[6] * inv, a[7] * inv);
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

extern "C" int tbn_avgpool_feat(const void* x, float* y, int n, int h, int w, int c, int y_ld, int freq_only,
                                int dtype, void* stream) {
    TBN_CHECK_ARG(x && y && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && y_ld % 4 == 0 && y_ld >= c,
                  "avgpool_feat: bad arguments");
    const long long total = (long long)n * (freq_only ? w : 1) * (c / 8);
    const int blocks = grid_for(total);
    cudaStream_t st = as_stream(stream);
    if (dtype == TBN_F32) {
        if (freq_only)
            avgpool_feat_kernel<float, true><<<blocks, 256, 0, st>>>((const float*)x, y, n, h, w, c, y_ld);
        else
            avgpool_feat_kernel<float, false><<<blocks, 256, 0, st>>>((const float*)x, y, n, h, w, c, y_ld);
    } else if (dtype == TBN_BF16) {
        if (freq_only)
            avgpool_feat_kernel<__nv_bfloat16, true><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, y, n, h, w, c, y_ld);
        else
            avgpool_feat_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, y, n, h, w, c, y_ld);
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
    const long long total = (long long)n * h * w * (c / 8);
    if (dtype == TBN_F32)
        avgpool3s1_bias_relu_kernel<float><<<grid_for(total), 256, 0, as_stream(stream)>>>((const float*)x, bias, (float*)y, g);
    else if (dtype == TBN_BF16)
        avgpool3s1_bias_relu_kernel<__nv_bfloat16>
            <<<grid_for(total), 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, bias, (__nv_bfloat16*)y, g);
    else
        TBN_CHECK_ARG(false, "avgpool3s1_bias_relu: bad dtype %d", dtype);
    TBN_LAUNCH_CHECK("avgpool3s1_bias_relu launch");
    return 0;
}
