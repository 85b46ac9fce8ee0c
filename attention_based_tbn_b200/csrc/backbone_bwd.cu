// Backward pass of the BN-Inception backbones (what torch autograd derives for BNInception.features + the feature pooling
// of reference bn_inception.py:16-35 under model.train(), core/tools/train.py:81), fp32 on CUDA cores.
//   wgrad  : fp32 im2col (tbn_im2col_f32; 1x1 convs read the activation directly) + split-K transposed GEMM
//            (tbn_sgemm_splitk: the reduction runs over all N*OH*OW output positions, the result is only K x Cout)
//   dgrad  : the forward conv kernel (tbn_conv_simt) on the output gradient with flipped / transposed weights; stride-2
//            convs first spread the gradient onto the input grid (tbn_upsample_zero2)
//   pools  : max-pool backward re-derives the arg-max (first maximum in scan order, as torch) and scatters; the 3x3/1
//            average pool is self-adjoint (tbn_pool3s1 on the gradient); feature pooling broadcasts
//   BatchNorm backward lives next to its forward in bn_train.cu.
// Activations are NHWC channel slices (cs = channel stride of the buffer, co = first channel).
#include <float.h>

#include "common.cuh"

namespace tbn {

__global__ void im2col_f32_kernel(const float* __restrict__ x, float* __restrict__ cols, int n, int h, int w, int cin, int x_cs, int x_co,
                                  int k, int stride, int pad, int oh, int ow) {
    const long long kk = (long long)k * k * cin;
    const long long total = (long long)n * oh * ow * kk;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(t % cin);
        long long r = t / cin;
        const int tap = (int)(r % (k * k));
        r /= (k * k);
        const int ox = (int)(r % ow);
        r /= ow;
        const int oy = (int)(r % oh);
        const int ni = (int)(r / oh);
        const int iy = oy * stride + tap / k - pad, ix = ox * stride + tap % k - pad;
        float v = 0.f;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = x[((long long)(ni * h + iy) * w + ix) * x_cs + x_co + ci];
        cols[t] = v;
    }
}

// the same, 4 input channels per thread (cin, x_cs, x_co multiples of 4)
__global__ void im2col_f32_v4_kernel(const float* __restrict__ x, float* __restrict__ cols, int n, int h, int w, int cin, int x_cs,
                                     int x_co, int k, int stride, int pad, int oh, int ow) {
    const int c4 = cin >> 2;
    const long long total = (long long)n * oh * ow * k * k * c4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(t % c4) * 4;
        long long r = t / c4;
        const int tap = (int)(r % (k * k));
        r /= (k * k);
        const int ox = (int)(r % ow);
        r /= ow;
        const int oy = (int)(r % oh);
        const int ni = (int)(r / oh);
        const int iy = oy * stride + tap / k - pad, ix = ox * stride + tap % k - pad;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < h && ix >= 0 && ix < w)
            v = *reinterpret_cast<const float4*>(x + ((long long)(ni * h + iy) * w + ix) * x_cs + x_co + ci);
        reinterpret_cast<float4*>(cols)[t] = v;
    }
}

// C (m x n) += A^T-or-A (m x k) * B (k x n) over the k range of blockIdx.z, atomically (C zeroed by the caller).
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_splitk_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B,
                                                           long long ldb, float* __restrict__ C, long long ldc, int m, int n,
                                                           long long k, long long kchunk, float alpha) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const long long kbeg = (long long)blockIdx.z * kchunk;
    const long long kend = kbeg + kchunk < k ? kbeg + kchunk : k;
    float acc[4][4] = {};
    for (long long k0 = kbeg; k0 < kend; k0 += 16) {
        for (int i = tid; i < 64 * 16; i += 256) {
            int r, kk;
            if (TA) { r = i & 63; kk = i >> 6; } else { kk = i & 15; r = i >> 4; }
            const int gr = m0 + r;
            const long long gk = k0 + kk;
            float v = 0.f;
            if (gr < m && gk < kend) v = TA ? A[gk * lda + gr] : A[(long long)gr * lda + gk];
            As[kk][r] = v;
        }
        for (int i = tid; i < 64 * 16; i += 256) {
            int c, kk;
            if (TB) { kk = i & 15; c = i >> 4; } else { c = i & 63; kk = i >> 6; }
            const int gc = n0 + c;
            const long long gk = k0 + kk;
            float v = 0.f;
            if (gc < n && gk < kend) v = TB ? B[(long long)gc * ldb + gk] : B[gk * ldb + gc];
            Bs[kk][c] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = m0 + ty * 4 + i;
        if (r >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            if (c < n) atomicAdd(C + (long long)r * ldc + c, alpha * acc[i][j]);
        }
    }
}

// The weight-gradient shape of the same product (A stored (k x m), B stored (k x n): both operands are read along their
// contiguous axis, rows of the reduction dimension): 128 x 64 tile, k step 16, 8 x 4 outputs per thread, the next k slab is
// fetched into registers while the current one is multiplied.
__global__ void __launch_bounds__(256) sgemm_tn_splitk_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B,
                                                              long long ldb, float* __restrict__ C, long long ldc, int m, int n,
                                                              long long k, long long kchunk, float alpha) {
    constexpr int BM = 128, BN = 64, BK = 16;
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // ty: 16 row groups of 8, tx: 16 column groups of 4
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const long long kbeg = (long long)blockIdx.z * kchunk;
    const long long kend = kbeg + kchunk < k ? kbeg + kchunk : k;
    // loader mapping: A slab 16 x 128 = 8 per thread (row kk = tid / 128 + 2 i, column tid % 128); B slab 16 x 64 = 4 per thread
    const int a_col = tid & 127, a_row = tid >> 7;
    const int b_col = tid & 63, b_row = tid >> 6;
    const bool a_ok = m0 + a_col < m, b_ok = n0 + b_col < n;
    float ra[8], rb[4];
    auto fetch = [&](long long k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long gk = k0 + a_row + 2 * i;
            ra[i] = (a_ok && gk < kend) ? A[gk * lda + m0 + a_col] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long gk = k0 + b_row + 4 * i;
            rb[i] = (b_ok && gk < kend) ? B[gk * ldb + n0 + b_col] : 0.f;
        }
    };
    float acc[8][4] = {};
    fetch(kbeg);
    for (long long k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[a_row + 2 * i][a_col] = ra[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) Bs[b_row + 4 * i][b_col] = rb[i];
        __syncthreads();
        if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (r >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            if (c < n) atomicAdd(C + (long long)r * ldc + c, alpha * acc[i][j]);
        }
    }
}

__global__ void zero_view_kernel(float* __restrict__ c, long long ldc, int m, int n) {
    const long long total = (long long)m * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
        c[(t / n) * ldc + t % n] = 0.f;
}

// y (n, H, W, c) = x (n, oh, ow, c) spread with stride 2 (zeros elsewhere): the input-grid image of a stride-2 conv's output gradient
__global__ void upsample_zero2_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int oh, int ow, int H, int W, int c) {
    const long long total = (long long)n * H * W * c;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(t % c);
        long long r = t / c;
        const int ix = (int)(r % W);
        r /= W;
        const int iy = (int)(r % H);
        const int ni = (int)(r / H);
        float v = 0.f;
        if (!(iy & 1) && !(ix & 1) && (iy >> 1) < oh && (ix >> 1) < ow) v = x[((long long)(ni * oh + (iy >> 1)) * ow + (ix >> 1)) * c + ci];
        y[t] = v;
    }
}

// Max-pool backward, 3x3 window.  S2: stride 2, ceil mode, no padding (windows may overhang: clipped); else stride 1, pad 1.
// dx[argmax] += dy, first maximum in (row, column) scan order.
template <bool S2>
__global__ void maxpool3_bwd_kernel(const float* __restrict__ x, int x_cs, int x_co, const float* __restrict__ dy, int dy_cs, int dy_co,
                                    float* __restrict__ dx, int dx_cs, int dx_co, int n, int h, int w, int c, int oh, int ow) {
    const long long total = (long long)n * oh * ow * c;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(t % c);
        long long r = t / c;
        const int ox = (int)(r % ow);
        r /= ow;
        const int oy = (int)(r % oh);
        const int ni = (int)(r / oh);
        const int y0 = S2 ? oy * 2 : oy - 1, x0 = S2 ? ox * 2 : ox - 1;
        float best = -FLT_MAX;
        int by = -1, bx = -1;
        for (int dy_ = 0; dy_ < 3; ++dy_) {
            const int iy = y0 + dy_;
            if (iy < 0 || iy >= h) continue;
            for (int dx_ = 0; dx_ < 3; ++dx_) {
                const int ix = x0 + dx_;
                if (ix < 0 || ix >= w) continue;
                const float v = x[((long long)(ni * h + iy) * w + ix) * x_cs + x_co + ci];
                if (v > best || by < 0) { best = v; by = iy; bx = ix; }
            }
        }
        const float g = dy[((long long)(ni * oh + oy) * ow + ox) * dy_cs + dy_co + ci];
        atomicAdd(dx + ((long long)(ni * h + by) * w + bx) * dx_cs + dx_co + ci, g);
    }
}

// dst[r][dco + j] += src[r][sco + j]
__global__ void add_view_kernel(float* __restrict__ dst, int d_cs, int d_co, const float* __restrict__ src, int s_cs, int s_co,
                                long long m, int c) {
    if (((c | d_cs | d_co | s_cs | s_co) & 3) == 0 && m * (c >> 2) < (1ll << 32)) {  // 4 channels per thread, 32-bit index math
        const unsigned c4 = (unsigned)c >> 2;
        const unsigned long long total = (unsigned long long)m * c4;
        for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
             t += (unsigned long long)gridDim.x * blockDim.x) {
            const unsigned t32 = (unsigned)t, r = t32 / c4, ci = (t32 - r * c4) * 4;
            float4* d = reinterpret_cast<float4*>(dst + (size_t)r * d_cs + d_co + ci);
            const float4 a = *d, b = *reinterpret_cast<const float4*>(src + (size_t)r * s_cs + s_co + ci);
            *d = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        }
        return;
    }
    const long long total = m * c;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(t % c);
        const long long r = t / c;
        dst[r * d_cs + d_co + ci] += src[r * s_cs + s_co + ci];
    }
}

// feature pooling backward (bn_inception.py:16-35): dx[n][y][x][c] = dfeat[n][c] / (h w)  or  dfeat[n*w + x][c] / h
__global__ void avgpool_feat_bwd_kernel(const float* __restrict__ df, long long ld, float* __restrict__ dx, int n, int h, int w, int c,
                                        int freq_only) {
    const long long total = (long long)n * h * w * c;
    const float inv = 1.0f / (freq_only ? h : h * w);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(t % c);
        long long r = t / c;
        const int ix = (int)(r % w);
        r /= w;
        const long long ni = r / h;
        const long long row = freq_only ? ni * w + ix : ni;
        dx[t] = df[row * ld + ci] * inv;
    }
}

// da[(f, s)][c] = wts[f][s] * dy[f][c]: gradient of feature[f] = sum_s a[f][s] w[f][s] w.r.t. the positions
__global__ void weighted_sum_bwd_a_kernel(const float* __restrict__ wts, const float* __restrict__ dy, long long ld_dy,
                                          float* __restrict__ da, int f, int s, int c) {
    const long long total = (long long)f * s * c;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(t % c);
        const long long fs = t / c;
        da[t] = wts[fs] * dy[(fs / s) * ld_dy + ci];
    }
}

// dst (m, c) bf16 contiguous = src[r][s_co + j] fp32: the tensor-core operand copy of an fp32 activation / gradient slice
__global__ void cast_bf16_view_kernel(const float* __restrict__ src, int s_cs, int s_co, __nv_bfloat16* __restrict__ dst, long long m,
                                      int c) {
    const unsigned c4 = (unsigned)c >> 2;
    const unsigned long long total = (unsigned long long)m * c4;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long r = t / c4;
        const unsigned ci = (unsigned)(t - r * c4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + r * s_cs + s_co + ci);
        uint2 o;
        *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + t * 4) = o;
    }
}

// Transposed bf16 im2col: dst[tap * cin + ci][p] = x[pixel p shifted by the tap][x_co + ci] (zero padding; zero for p >= M),
// row length mpad -- the reduction-major operand the tensor-core weight-gradient GEMM wants.  With k = 1 this is a plain
// transpose-and-cast (the output gradient du, or the input of a 1x1 conv).  One block = 64 pixels x 32 channels, all taps:
// the pixel coordinates are decoded once, every tap is a coalesced 128 B-row read, a transpose through shared memory and
// 32 coalesced 128 B row segments out.
__global__ void __launch_bounds__(256) im2col_t_bf16_kernel(const float* __restrict__ x, int x_cs, int x_co, __nv_bfloat16* __restrict__ dst,
                                                            long long mpad, int h, int w, int cin, int k, int stride, int pad, int oh,
                                                            int ow, long long M, int tiles_per_block) {
    __shared__ float tile[64][33];
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // load role: channel tx, pixels ty + 8 i
    const int px2 = threadIdx.x & 31, cy = threadIdx.x >> 5;  // store role: pixels 2 px2, 2 px2 + 1 (one 4-byte store), channels cy + 8 i
    const bool c_ok = c0 + tx < cin;
    const int taps = k * k;
    for (int tb = 0; tb < tiles_per_block; ++tb) {
        const long long p0 = ((long long)blockIdx.x * tiles_per_block + tb) * 64;
        if (p0 >= mpad) break;
        long long base[8];  // element offset of image ni in x, or -1 for p >= M
        int iy0[8], ix0[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long p = p0 + ty + 8 * i;
            base[i] = -1;
            iy0[i] = ix0[i] = 0;
            if (p < M) {
                const long long ni = p / ((long long)oh * ow);
                const int r = (int)(p - ni * oh * ow);
                const int oy = r / ow, ox = r - oy * ow;
                iy0[i] = oy * stride - pad;
                ix0[i] = ox * stride - pad;
                base[i] = ni * h * w;
            }
        }
        for (int tap = 0; tap < taps; ++tap) {
            const int ky = tap / k, kx = tap - ky * k;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = 0.f;
                const int iy = iy0[i] + ky, ix = ix0[i] + kx;
                if (c_ok && base[i] >= 0 && iy >= 0 && iy < h && ix >= 0 && ix < w)
                    v = x[(base[i] + (long long)iy * w + ix) * x_cs + x_co + c0 + tx];
                tile[ty + 8 * i][tx] = v;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cl = cy + 8 * i;
                if (c0 + cl < cin)  // (mpad is a multiple of 64: the whole pixel tile is inside the row)
                    *reinterpret_cast<__nv_bfloat162*>(dst + ((long long)tap * cin + c0 + cl) * mpad + p0 + 2 * px2) =
                        __floats2bfloat162_rn(tile[2 * px2][cl], tile[2 * px2 + 1][cl]);
            }
            __syncthreads();
        }
    }
}

static inline unsigned grid_for(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_im2col_f32(const float* x, float* cols, int n, int h, int w, int cin, int x_cs, int x_co, int k, int stride, int pad,
                              void* stream) {
    TBN_CHECK_ARG(x && cols && n > 0 && cin > 0 && k > 0 && stride > 0, "im2col_f32: bad arguments");
    const int oh = (h + 2 * pad - k) / stride + 1, ow = (w + 2 * pad - k) / stride + 1;
    const long long total = (long long)n * oh * ow * k * k * cin;
    if (cin % 4 == 0 && x_cs % 4 == 0 && x_co % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)cols & 15) == 0)
        im2col_f32_v4_kernel<<<grid_for(total / 4), 256, 0, as_stream(stream)>>>(x, cols, n, h, w, cin, x_cs, x_co, k, stride, pad, oh, ow);
    else
        im2col_f32_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, cols, n, h, w, cin, x_cs, x_co, k, stride, pad, oh, ow);
    TBN_LAUNCH_CHECK("im2col_f32 launch");
    return 0;
}

extern "C" int tbn_sgemm_splitk(int ta, int tb, int m, int n, long long k, float alpha, const float* a, long long lda, const float* b,
                                long long ldb, float* c, long long ldc, int ksplit, void* stream) {
    TBN_CHECK_ARG(a && b && c && m > 0 && n > 0 && k > 0 && ksplit >= 1, "sgemm_splitk: bad arguments");
    cudaStream_t st = as_stream(stream);
    zero_view_kernel<<<grid_for((long long)m * n), 256, 0, st>>>(c, ldc, m, n);
    TBN_LAUNCH_CHECK("sgemm_splitk zero");
    long long kchunk = (k + ksplit - 1) / ksplit;
    kchunk = (kchunk + 15) / 16 * 16;
    const int nz = (int)((k + kchunk - 1) / kchunk);
    if (ta && !tb) {
        dim3 g2(cdiv(n, 64), cdiv(m, 128), nz);
        sgemm_tn_splitk_kernel<<<g2, 256, 0, st>>>(a, lda, b, ldb, c, ldc, m, n, k, kchunk, alpha);
        TBN_LAUNCH_CHECK("sgemm_tn_splitk launch");
        return 0;
    }
    dim3 grid(cdiv(n, 64), cdiv(m, 64), nz);
    if (ta && tb) sgemm_splitk_kernel<true, true><<<grid, 256, 0, st>>>(a, lda, b, ldb, c, ldc, m, n, k, kchunk, alpha);
    else if (ta) sgemm_splitk_kernel<true, false><<<grid, 256, 0, st>>>(a, lda, b, ldb, c, ldc, m, n, k, kchunk, alpha);
    else if (tb) sgemm_splitk_kernel<false, true><<<grid, 256, 0, st>>>(a, lda, b, ldb, c, ldc, m, n, k, kchunk, alpha);
    else sgemm_splitk_kernel<false, false><<<grid, 256, 0, st>>>(a, lda, b, ldb, c, ldc, m, n, k, kchunk, alpha);
    TBN_LAUNCH_CHECK("sgemm_splitk launch");
    return 0;
}

extern "C" int tbn_upsample_zero2(const float* x, float* y, int n, int oh, int ow, int H, int W, int c, void* stream) {
    TBN_CHECK_ARG(x && y && n > 0 && oh > 0 && ow > 0 && 2 * (oh - 1) < H && 2 * (ow - 1) < W, "upsample_zero2: bad arguments");
    upsample_zero2_kernel<<<grid_for((long long)n * H * W * c), 256, 0, as_stream(stream)>>>(x, y, n, oh, ow, H, W, c);
    TBN_LAUNCH_CHECK("upsample_zero2 launch");
    return 0;
}

extern "C" int tbn_maxpool3_bwd(const float* x, int x_cs, int x_co, const float* dy, int dy_cs, int dy_co, float* dx, int dx_cs, int dx_co,
                                int n, int h, int w, int c, int stride2, void* stream) {
    TBN_CHECK_ARG(x && dy && dx && n > 0 && c > 0, "maxpool3_bwd: bad arguments");
    int oh = h, ow = w;
    if (stride2) {
        oh = (h - 3 + 1) / 2 + 1;
        ow = (w - 3 + 1) / 2 + 1;
        if ((oh - 1) * 2 >= h) --oh;
        if ((ow - 1) * 2 >= w) --ow;
    }
    const long long total = (long long)n * oh * ow * c;
    if (stride2)
        maxpool3_bwd_kernel<true><<<grid_for(total), 256, 0, as_stream(stream)>>>(x, x_cs, x_co, dy, dy_cs, dy_co, dx, dx_cs, dx_co, n, h, w, c,
                                                                                   oh, ow);
    else
        maxpool3_bwd_kernel<false><<<grid_for(total), 256, 0, as_stream(stream)>>>(x, x_cs, x_co, dy, dy_cs, dy_co, dx, dx_cs, dx_co, n, h, w,
                                                                                    c, oh, ow);
    TBN_LAUNCH_CHECK("maxpool3_bwd launch");
    return 0;
}

extern "C" int tbn_add_view(float* dst, int d_cs, int d_co, const float* src, int s_cs, int s_co, long long m, int c, void* stream) {
    TBN_CHECK_ARG(dst && src && m > 0 && c > 0 && d_co + c <= d_cs && s_co + c <= s_cs, "add_view: bad arguments");
    add_view_kernel<<<grid_for(m * c), 256, 0, as_stream(stream)>>>(dst, d_cs, d_co, src, s_cs, s_co, m, c);
    TBN_LAUNCH_CHECK("add_view launch");
    return 0;
}

extern "C" int tbn_avgpool_feat_bwd(const float* dfeat, long long ld, float* dx, int n, int h, int w, int c, int freq_only, void* stream) {
    TBN_CHECK_ARG(dfeat && dx && n > 0 && h > 0 && w > 0 && c > 0 && ld >= c, "avgpool_feat_bwd: bad arguments");
    avgpool_feat_bwd_kernel<<<grid_for((long long)n * h * w * c), 256, 0, as_stream(stream)>>>(dfeat, ld, dx, n, h, w, c, freq_only);
    TBN_LAUNCH_CHECK("avgpool_feat_bwd launch");
    return 0;
}

extern "C" int tbn_weighted_sum_bwd_a(const float* wts, const float* dy, long long ld_dy, float* da, int f, int s, int c, void* stream) {
    TBN_CHECK_ARG(wts && dy && da && f > 0 && s > 0 && c > 0, "weighted_sum_bwd_a: bad arguments");
    weighted_sum_bwd_a_kernel<<<grid_for((long long)f * s * c), 256, 0, as_stream(stream)>>>(wts, dy, ld_dy, da, f, s, c);
    TBN_LAUNCH_CHECK("weighted_sum_bwd_a launch");
    return 0;
}

extern "C" int tbn_cast_bf16_view(const float* src, int s_cs, int s_co, void* dst, long long m, int c, void* stream) {
    TBN_CHECK_ARG(src && dst && m > 0 && c > 0 && ((c | s_cs | s_co) & 3) == 0 && s_co + c <= s_cs, "cast_bf16_view: bad arguments");
    cast_bf16_view_kernel<<<grid_for(m * (c / 4)), 256, 0, as_stream(stream)>>>(src, s_cs, s_co, reinterpret_cast<__nv_bfloat16*>(dst), m, c);
    TBN_LAUNCH_CHECK("cast_bf16_view launch");
    return 0;
}

extern "C" int tbn_im2col_t_bf16(const float* x, int x_cs, int x_co, void* dst, long long mpad, int n, int h, int w, int cin, int k,
                                 int stride, int pad, void* stream) {
    TBN_CHECK_ARG(x && dst && n > 0 && cin > 0 && k > 0 && stride > 0, "im2col_t_bf16: bad arguments");
    const int oh = (h + 2 * pad - k) / stride + 1, ow = (w + 2 * pad - k) / stride + 1;
    const long long M = (long long)n * oh * ow;
    TBN_CHECK_ARG(mpad >= M && mpad % 64 == 0 && mpad / 64 < (1ll << 31), "im2col_t_bf16: the row length must be a multiple of 64 >= n*oh*ow");
    TBN_CHECK_ARG(((uintptr_t)dst & 3) == 0, "im2col_t_bf16: dst must be 4-byte aligned");
    // few taps = little work per 64-pixel tile: let a block walk several tiles (amortises launch + coordinate decoding)
    const int tiles_per_block = k == 1 ? 8 : (k == 3 ? 2 : 1);
    const long long ptiles = mpad / 64;
    dim3 grid((unsigned)((ptiles + tiles_per_block - 1) / tiles_per_block), (unsigned)((cin + 31) / 32));
    im2col_t_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_cs, x_co, reinterpret_cast<__nv_bfloat16*>(dst), mpad, h, w, cin, k, stride,
                                                              pad, oh, ow, M, tiles_per_block);
    TBN_LAUNCH_CHECK("im2col_t_bf16 launch");
    return 0;
}
