// Backward pass of the TBN heads (reference model.py:214-262 under autograd: Classifier, Fusion, nn.MultiheadAttention,
// pe = PositionalEncoding -> Conv1d -> GroupNorm, and the MLP attentions of attention.py:60-145).  fp32 CUDA-core kernels
// -- the reference trains in fp32 and the heads are ~1 % of a training step's arithmetic.  The backbones stay frozen
// (the reference's freeze_base / freeze_mode = "all" set-up, model.py:160-163); their backward is not part of this file.
//
// Everything is row-major with explicit leading dimensions; rows = frames (F) or frame-positions (F*S).
#include "common.cuh"

namespace tbn {

// ---------------------------------------------------------------------------------------------
// C (m x n) = alpha * op(A) (m x k) * op(B) (k x n) + beta * C.   ta: A is stored (k x m); tb: B is stored (n x k).
// 64 x 64 tile, k step 16, 256 threads, 4 x 4 outputs per thread.
// ---------------------------------------------------------------------------------------------
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B,
                                                    long long ldb, float* __restrict__ C, long long ldc, int m, int n, int k,
                                                    float alpha, float beta) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < k; k0 += 16) {
        // 64 x 16 elements of each operand, 4 per thread; the fast thread index runs along the operand's contiguous axis
        for (int i = tid; i < 64 * 16; i += 256) {
            int r, kk;
            if (TA) { r = i & 63; kk = i >> 6; } else { kk = i & 15; r = i >> 4; }
            const int gr = m0 + r, gk = k0 + kk;
            float v = 0.f;
            if (gr < m && gk < k) v = TA ? A[(long long)gk * lda + gr] : A[(long long)gr * lda + gk];
            As[kk][r] = v;
        }
        for (int i = tid; i < 64 * 16; i += 256) {
            int c, kk;
            if (TB) { kk = i & 15; c = i >> 4; } else { c = i & 63; kk = i >> 6; }
            const int gc = n0 + c, gk = k0 + kk;
            float v = 0.f;
            if (gc < n && gk < k) v = TB ? B[(long long)gc * ldb + gk] : B[(long long)gk * ldb + gc];
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
            if (c >= n) continue;
            float* p = C + (long long)r * ldc + c;
            *p = beta == 0.f ? alpha * acc[i][j] : alpha * acc[i][j] + beta * *p;
        }
    }
}

// out[j] = beta * out[j] + sum_i x[i][j]   (bias gradients); one warp-wide column strip per block, fixed summation order
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long ld, int m, int n, float* __restrict__ out,
                                                     float beta) {
    __shared__ float part[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float acc = 0.f;
    if (c < n)
        for (int i = w; i < m; i += 8) acc += x[(long long)i * ld + c];
    part[w][lane] = acc;
    __syncthreads();
    if (w == 0 && c < n) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][lane];
        out[c] = beta == 0.f ? t : t + beta * out[c];
    }
}

// segment consensus backward: dlogits[b*nseg + s][c] = dscores[b][c] / nseg   (reference model.py:194-199)
__global__ void consensus_bwd_kernel(const float* __restrict__ ds, long long ld_ds, float* __restrict__ dl, int b, int nseg, int c) {
    const long long total = (long long)b * nseg * c;
    const float inv = 1.0f / nseg;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(t % c);
        const long long row = t / c;
        dl[t] = ds[(row / nseg) * ld_ds + ci] * inv;
    }
}

// ReLU (+ inverted dropout) backward from the layer's OUTPUT h: dh *= (h > 0) * scale.  A dropped or clamped unit has
// h == 0 and passes no gradient either way; scale = 1 / (1 - p) of the dropout that followed the ReLU (1 without).
__global__ void relu_dropout_bwd_kernel(float* __restrict__ dh, const float* __restrict__ h, float scale, long long n) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        dh[t] = h[t] > 0.f ? dh[t] * scale : 0.f;
}

// o[f][h*d + j] += bv[h*d + j] * (psum[f][h] - 1): with dropout on the attention probabilities the value bias is weighted by
// the sum of the kept probabilities, not by 1 (o_h = sum_s p_s (W_v z_s + b_v))
__global__ void value_bias_fix_kernel(float* __restrict__ o, const float* __restrict__ bv, const float* __restrict__ psum, int f, int e,
                                      int heads) {
    const long long total = (long long)f * e;
    const int d = e / heads;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(t % e);
        const long long fi = t / e;
        o[t] += bv[c] * (psum[fi * heads + c / d] - 1.f);
    }
}

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// GroupNorm backward over rows (f, s) x channels: x = pre-norm input (F*S, C), dy = gradient of the output.  One warp per
// (frame, group): dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = gamma * dy; dgamma / dbeta are written per frame
// into (F, C) partial tables (reduced by colsum_kernel: deterministic, no atomics).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) groupnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            const float* __restrict__ gamma, float* __restrict__ dx,
                                                            float* __restrict__ dgamma_f, float* __restrict__ dbeta_f, int f, int s,
                                                            int c, int groups, float eps) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= f * groups) return;
    const int fi = gw / groups, g = gw - fi * groups;
    const int cpg = c / groups, cnt = cpg * s;
    const long long base = (long long)fi * s * c + g * cpg;
    float sum = 0.f;
    for (int t = lane; t < cnt; t += 32) sum += x[base + (long long)(t / cpg) * c + (t % cpg)];
    const float mean = wsum(sum) / cnt;
    float sq = 0.f;
    for (int t = lane; t < cnt; t += 32) {
        const float d = x[base + (long long)(t / cpg) * c + (t % cpg)] - mean;
        sq += d * d;
    }
    const float rstd = rsqrtf(wsum(sq) / cnt + eps);
    float sg = 0.f, sgx = 0.f;
    for (int t = lane; t < cnt; t += 32) {
        const long long o = base + (long long)(t / cpg) * c + (t % cpg);
        const float gg = gamma[g * cpg + t % cpg] * dy[o];
        sg += gg;
        sgx += gg * (x[o] - mean) * rstd;
    }
    const float mg = wsum(sg) / cnt, mgx = wsum(sgx) / cnt;
    for (int t = lane; t < cnt; t += 32) {
        const long long o = base + (long long)(t / cpg) * c + (t % cpg);
        const float xh = (x[o] - mean) * rstd;
        dx[o] = rstd * (gamma[g * cpg + t % cpg] * dy[o] - mg - xh * mgx);
    }
    // per-channel partials of this frame: sum over positions
    for (int ci = lane; ci < cpg; ci += 32) {
        float a = 0.f, b = 0.f;
        for (int p = 0; p < s; ++p) {
            const long long o = base + (long long)p * c + ci;
            a += dy[o] * (x[o] - mean) * rstd;
            b += dy[o];
        }
        dgamma_f[(long long)fi * c + g * cpg + ci] = a;
        dbeta_f[(long long)fi * c + g * cpg + ci] = b;
    }
}

// ---------------------------------------------------------------------------------------------
// Single-query multi-head attention backward (torch.nn.functional.multi_head_attention_forward with L = 1):
//   p = softmax_s(q_h . k_{s,h} / sqrt(d)),  pt = p * mask (inverted dropout),  o_h = sum_s pt_s v_{s,h},  w = mean_h pt.
// One CTA per frame, one warp per head (loops when heads > 8).  Inputs q (F, E), k, v (F*S, E), mask (F, H, S) or null,
// d_o (F, E), d_w (F, S) or null.  Outputs dq (F, E), dk, dv (F*S, E).
// ---------------------------------------------------------------------------------------------
constexpr int BWD_MAX_S = 32;

__global__ void __launch_bounds__(256) mha_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                      const float* __restrict__ mask, const float* __restrict__ d_o,
                                                      const float* __restrict__ d_w, float* __restrict__ dq, float* __restrict__ dk,
                                                      float* __restrict__ dv, int s, int e, int heads) {
    __shared__ float s_ds[8][32], s_pt[8][32];
    const int f = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = e / heads;
    const float scale = rsqrtf((float)d);
    for (int h = warp; h < heads; h += 8) {
        const float* qh = q + (long long)f * e + h * d;
        const float* doh = d_o + (long long)f * e + h * d;
        const float* kf = k + (long long)f * s * e + h * d;
        const float* vf = v + (long long)f * s * e + h * d;
        // lane p keeps position p's score and d(pt)
        float sc = -INFINITY, dpt = 0.f;
        for (int p = 0; p < s; ++p) {
            float a = 0.f, b = 0.f;
            for (int j = lane; j < d; j += 32) {
                a = fmaf(qh[j], kf[(long long)p * e + j], a);
                b = fmaf(doh[j], vf[(long long)p * e + j], b);
            }
            a = wsum(a) * scale;
            b = wsum(b);
            if (lane == p) { sc = a; dpt = b; }
        }
        const float mx = wmax(sc);
        const float ex = lane < s ? expf(sc - mx) : 0.f;
        const float pr = ex / wsum(ex);
        const float keep = (mask != nullptr && lane < s) ? mask[((long long)f * heads + h) * s + lane] : 1.f;
        const float pt = pr * keep;
        if (d_w != nullptr && lane < s) dpt += d_w[(long long)f * s + lane] / heads;
        const float dp = lane < s ? dpt * keep : 0.f;
        const float dot = wsum(pr * dp);
        const float ds = lane < s ? pr * (dp - dot) * scale : 0.f;  // d(q.k), 1/sqrt(d) included
        __syncwarp();
        s_ds[warp][lane] = ds;
        s_pt[warp][lane] = pt;
        __syncwarp();
        for (int j = lane; j < d; j += 32) {
            float aq = 0.f;
            const float qj = qh[j], dj = doh[j];
            for (int p = 0; p < s; ++p) {
                const float dsp = s_ds[warp][p], ptp = s_pt[warp][p];
                aq = fmaf(dsp, kf[(long long)p * e + j], aq);
                dk[((long long)f * s + p) * e + h * d + j] = dsp * qj;
                dv[((long long)f * s + p) * e + h * d + j] = ptp * dj;
            }
            dq[(long long)f * e + h * d + j] = aq;
        }
    }
}

// Weighted-sum attention backward (fixed / MLP attentions: feature[f][c] = sum_s a[f][s][c] w[f][s]):
// dw[f][s] = (dw_in ? dw_in[f][s] : 0) + sum_c dy[f][c] a[f][s][c];  one warp per (f, s)
__global__ void __launch_bounds__(256) weighted_sum_bwd_kernel(const float* __restrict__ a, const float* __restrict__ dy, long long ld_dy,
                                                               const float* __restrict__ dw_in, float* __restrict__ dw, int f, int s,
                                                               int c) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= f * s) return;
    const int fi = gw / s;
    const float* ar = a + (long long)gw * c;
    const float* dr = dy + (long long)fi * ld_dy;
    float acc = 0.f;
    for (int j = lane; j < c; j += 32) acc = fmaf(ar[j], dr[j], acc);
    acc = wsum(acc);
    if (lane == 0) dw[gw] = acc + (dw_in ? dw_in[gw] : 0.f);
}

// softmax backward on rows: dx = y * (dy - sum(y * dy)); one warp per row, cols <= 1024
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx,
                                                          int rows, int cols, long long ld) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    float dot = 0.f;
    for (int i = lane; i < cols; i += 32) dot += y[row * ld + i] * dy[row * ld + i];
    dot = wsum(dot);
    for (int i = lane; i < cols; i += 32) dx[row * ld + i] = y[row * ld + i] * (dy[row * ld + i] - dot);
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_sgemm(int ta, int tb, int m, int n, int k, float alpha, const float* a, long long lda, const float* b, long long ldb,
                         float beta, float* c, long long ldc, void* stream) {
    TBN_CHECK_ARG(a && b && c && m > 0 && n > 0 && k > 0, "sgemm: null pointer or empty problem (m=%d n=%d k=%d)", m, n, k);
    TBN_CHECK_ARG(lda >= (ta ? m : k) && ldb >= (tb ? k : n) && ldc >= n, "sgemm: leading dimension smaller than the row length");
    if (ta && !tb && beta == 0.f && (long long)m * n >= 64 * 64) {
        // weight-gradient shape (both operands read along their contiguous axis): the register-tiled split-K kernel of
        // backbone_bwd.cu, with enough k splits to fill the GPU
        const long long tiles = (long long)cdiv(m, 128) * cdiv(n, 64);
        long long ksplit = (2ll * sm_count() + tiles - 1) / tiles;
        if (ksplit > k / 64) ksplit = k / 64;
        if (ksplit < 1) ksplit = 1;
        return tbn_sgemm_splitk(ta, tb, m, n, k, alpha, a, lda, b, ldb, c, ldc, (int)ksplit, stream);
    }
    dim3 grid(cdiv(n, 64), cdiv(m, 64));
    cudaStream_t st = as_stream(stream);
    if (ta && tb) sgemm_kernel<true, true><<<grid, 256, 0, st>>>(a, lda, b, ldb, c, ldc, m, n, k, alpha, beta);
    else if (ta) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(a, lda, b, ldb, c, ldc, m, n, k, alpha, beta);
    else if (tb) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(a, lda, b, ldb, c, ldc, m, n, k, alpha, beta);
    else sgemm_kernel<false, false><<<grid, 256, 0, st>>>(a, lda, b, ldb, c, ldc, m, n, k, alpha, beta);
    TBN_LAUNCH_CHECK("sgemm launch");
    return 0;
}

extern "C" int tbn_colsum(const float* x, long long ld, int m, int n, float beta, float* out, void* stream) {
    TBN_CHECK_ARG(x && out && m > 0 && n > 0 && ld >= n, "colsum: bad arguments");
    colsum_kernel<<<cdiv(n, 32), 256, 0, as_stream(stream)>>>(x, ld, m, n, out, beta);
    TBN_LAUNCH_CHECK("colsum launch");
    return 0;
}

extern "C" int tbn_consensus_bwd(const float* dscores, long long ld, float* dlogits, int b, int nseg, int c, void* stream) {
    TBN_CHECK_ARG(dscores && dlogits && b > 0 && nseg > 0 && c > 0 && ld >= c, "consensus_bwd: bad arguments");
    const long long total = (long long)b * nseg * c;
    consensus_bwd_kernel<<<(int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0, as_stream(stream)>>>(dscores, ld, dlogits, b,
                                                                                                                         nseg, c);
    TBN_LAUNCH_CHECK("consensus_bwd launch");
    return 0;
}

extern "C" int tbn_relu_dropout_bwd(float* dh, const float* h, float scale, long long n, void* stream) {
    TBN_CHECK_ARG(dh && h && n > 0, "relu_dropout_bwd: bad arguments");
    relu_dropout_bwd_kernel<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, as_stream(stream)>>>(dh, h, scale, n);
    TBN_LAUNCH_CHECK("relu_dropout_bwd launch");
    return 0;
}

extern "C" int tbn_value_bias_fix(float* o, const float* bv, const float* psum, int f, int e, int heads, void* stream) {
    TBN_CHECK_ARG(o && bv && psum && f > 0 && heads > 0 && e % heads == 0, "value_bias_fix: bad arguments");
    const long long n = (long long)f * e;
    value_bias_fix_kernel<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, as_stream(stream)>>>(o, bv, psum, f, e, heads);
    TBN_LAUNCH_CHECK("value_bias_fix launch");
    return 0;
}

extern "C" int tbn_groupnorm_bwd(const float* x, const float* dy, const float* gamma, float* dx, float* dgamma_f, float* dbeta_f, int f,
                                 int s, int c, int groups, float eps, void* stream) {
    TBN_CHECK_ARG(x && dy && gamma && dx && dgamma_f && dbeta_f && f > 0 && s > 0 && groups > 0 && c % groups == 0,
                  "groupnorm_bwd: bad arguments");
    const long long warps = (long long)f * groups;
    groupnorm_bwd_kernel<<<cdiv(warps * 32, 256), 256, 0, as_stream(stream)>>>(x, dy, gamma, dx, dgamma_f, dbeta_f, f, s, c, groups, eps);
    TBN_LAUNCH_CHECK("groupnorm_bwd launch");
    return 0;
}

extern "C" int tbn_mha_bwd(const float* q, const float* k, const float* v, const float* mask, const float* d_o, const float* d_w,
                           float* dq, float* dk, float* dv, int f, int s, int e, int heads, void* stream) {
    TBN_CHECK_ARG(q && k && v && d_o && dq && dk && dv && f > 0, "mha_bwd: null pointer");
    TBN_CHECK_ARG(s >= 1 && s <= BWD_MAX_S && heads >= 1 && e % heads == 0, "mha_bwd: need 1<=S<=32 and E %% heads == 0 (S=%d heads=%d E=%d)",
                  s, heads, e);
    mha_bwd_kernel<<<f, 256, 0, as_stream(stream)>>>(q, k, v, mask, d_o, d_w, dq, dk, dv, s, e, heads);
    TBN_LAUNCH_CHECK("mha_bwd launch");
    return 0;
}

extern "C" int tbn_weighted_sum_bwd(const float* a, const float* dy, long long ld_dy, const float* dw_in, float* dw, int f, int s, int c,
                                    void* stream) {
    TBN_CHECK_ARG(a && dy && dw && f > 0 && s > 0 && c > 0, "weighted_sum_bwd: bad arguments");
    weighted_sum_bwd_kernel<<<cdiv((long long)f * s * 32, 256), 256, 0, as_stream(stream)>>>(a, dy, ld_dy, dw_in, dw, f, s, c);
    TBN_LAUNCH_CHECK("weighted_sum_bwd launch");
    return 0;
}

extern "C" int tbn_softmax_bwd(const float* y, const float* dy, float* dx, int rows, int cols, long long ld, void* stream) {
    TBN_CHECK_ARG(y && dy && dx && rows > 0 && cols > 0 && ld >= cols, "softmax_bwd: bad arguments");
    softmax_bwd_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, as_stream(stream)>>>(y, dy, dx, rows, cols, ld);
    TBN_LAUNCH_CHECK("softmax_bwd launch");
    return 0;
}
