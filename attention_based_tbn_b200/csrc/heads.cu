// Head kernels (fp32): linear layers, GroupNorm, fused single-query multi-head attention core
// (scores + softmax + head-mean weights + entropy + weighted value sum), fixed-weight attention,
// row softmax and segment consensus.  All are small next to the backbones (<1 % of the FLOPs).
#include "common.cuh"

namespace tbn {

// ---------------------------------------------------------------------------------------------
// y[m][j] = act(sum_k x[m][k] * wt[j][k] + bias[(m % period)][j]); both operands K-contiguous.
// 256 threads; tile BMxBN, each thread TMxTN outputs; optional groups on blockIdx.z.
// ---------------------------------------------------------------------------------------------
struct LinearArgs {
    const float* x;
    const float* wt;
    const float* bias;
    float* y;
    int m, n, k, ldx, ldw, ldy, bias_period, relu;
    long long gx, gw, gb, gy;  // per-group element strides
};

template <int TBM, int TBN, int TM, int TN>
__global__ void __launch_bounds__(256) linear_kernel(LinearArgs a) {
    constexpr int TBK = 16;
    static_assert((TBM / TM) * (TBN / TN) == 256, "256 threads");
    __shared__ __align__(16) float As[TBK][TBM + 4];
    __shared__ __align__(16) float Bs[TBK][TBN + 4];
    const int tid = threadIdx.x;
    const int grp = blockIdx.z;
    const float* x = a.x + grp * a.gx;
    const float* wt = a.wt + grp * a.gw;
    const float* bias = a.bias ? a.bias + grp * a.gb : nullptr;
    float* y = a.y + grp * a.gy;
    const int m0 = blockIdx.x * TBM, n0 = blockIdx.y * TBN;
    const int tx = tid % (TBN / TN), ty = tid / (TBN / TN);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.k; k0 += TBK) {
        __syncthreads();
        // each operand tile is rows x 16 floats = rows*4 float4
        for (int v = tid; v < TBM * 4; v += 256) {
            const int row = v >> 2, kq = (v & 3) * 4;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + row < a.m && k0 + kq < a.k)
                t = *reinterpret_cast<const float4*>(x + (long long)(m0 + row) * a.ldx + k0 + kq);
            As[kq + 0][row] = t.x; As[kq + 1][row] = t.y; As[kq + 2][row] = t.z; As[kq + 3][row] = t.w;
        }
        for (int v = tid; v < TBN * 4; v += 256) {
            const int row = v >> 2, kq = (v & 3) * 4;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + row < a.n && k0 + kq < a.k)
                t = *reinterpret_cast<const float4*>(wt + (long long)(n0 + row) * a.ldw + k0 + kq);
            Bs[kq + 0][row] = t.x; Bs[kq + 1][row] = t.y; Bs[kq + 2][row] = t.z; Bs[kq + 3][row] = t.w;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < TBK; ++q) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) av[i] = As[q][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = Bs[q][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int mo = m0 + ty * TM + i;
        if (mo >= a.m) continue;
        const float* brow = bias ? bias + (long long)(mo % a.bias_period) * a.n : nullptr;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int no = n0 + tx * TN + j;
            if (no >= a.n) continue;
            float v = acc[i][j] + (brow ? brow[no] : 0.f);
            if (a.relu) v = fmaxf(v, 0.f);
            y[(long long)mo * a.ldy + no] = v;
        }
    }
}

static int launch_linear(const LinearArgs& a, int groups, cudaStream_t st) {
    TBN_CHECK_ARG(a.x && a.wt && a.y && a.m > 0 && a.n > 0 && a.k > 0, "linear: bad arguments");
    TBN_CHECK_ARG(a.k % 4 == 0 && a.ldx % 4 == 0 && a.ldw % 4 == 0 && a.gx % 4 == 0 && a.gw % 4 == 0,
                  "linear: k, ldx, ldw and group strides must be multiples of 4 floats");
    TBN_CHECK_ARG(a.bias_period >= 1, "linear: bias_period must be >= 1");
    const long long big_ctas = (long long)cdiv(a.m, 64) * cdiv(a.n, 64) * groups;
    if (big_ctas >= 2LL * sm_count()) {
        dim3 grid(cdiv(a.m, 64), cdiv(a.n, 64), groups);
        linear_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(a);
    } else {
        dim3 grid(cdiv(a.m, 32), cdiv(a.n, 32), groups);
        linear_kernel<32, 32, 2, 2><<<grid, 256, 0, st>>>(a);
    }
    TBN_LAUNCH_CHECK("linear launch");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// GroupNorm over position-major z (f, s, c): one warp per (sample, group).
// ---------------------------------------------------------------------------------------------
__global__ void groupnorm_kernel(float* __restrict__ z, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int f, int s, int c, int groups, float eps) {
    const int cpg = c / groups;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= f * groups) return;
    const int fi = warp / groups, g = warp - fi * groups;
    float* base = z + (long long)fi * s * c + g * cpg;
    const int cnt = cpg * s;
    float sum = 0.f;
    for (int e = lane; e < cnt; e += 32) sum += base[(long long)(e / cpg) * c + (e % cpg)];
    const float mean = warp_sum(sum) / cnt;
    float sq = 0.f;
    for (int e = lane; e < cnt; e += 32) {
        const float d = base[(long long)(e / cpg) * c + (e % cpg)] - mean;
        sq += d * d;
    }
    const float rstd = rsqrtf(warp_sum(sq) / cnt + eps);
    for (int e = lane; e < cnt; e += 32) {
        const int ci = e % cpg;
        float* p = base + (long long)(e / cpg) * c + ci;
        *p = (*p - mean) * rstd * gamma[g * cpg + ci] + beta[g * cpg + ci];
    }
}

// ---------------------------------------------------------------------------------------------
// Fused attention core: one CTA per frame, z[f] (S x E) and qk[f] (H x E) staged in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int MHA_MAX_S = 32, MHA_MAX_H = 8;

// gamma != nullptr: GroupNorm(groups, e) of z[f] is applied in shared memory first (same summation order as
// groupnorm_kernel: bit-identical to running it as a separate launch); zbar3 != nullptr: the weighted value sums are
// also written as the bf16 [hi | hi | lo] split operand of the out-projection GEMM (per head block of e columns).
__global__ void __launch_bounds__(256) mha_core_kernel(const float* __restrict__ qk, const float* __restrict__ z,
                                                       float* __restrict__ zbar, float* __restrict__ wts,
                                                       float* __restrict__ entropy, int s, int e, int heads,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int groups, float eps, __nv_bfloat16* __restrict__ zbar3,
                                                       const float* __restrict__ mask, float* __restrict__ psum) {
    extern __shared__ __align__(16) float smem[];
    float* zs = smem;                 // s*e
    float* qs = zs + (size_t)s * e;   // heads*e
    __shared__ float sc[MHA_MAX_H][MHA_MAX_S];
    const int f = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const float4* zg = reinterpret_cast<const float4*>(z + (long long)f * s * e);
    const float4* qg = reinterpret_cast<const float4*>(qk + (long long)f * heads * e);
    for (int i = tid; i < s * e / 4; i += 256) reinterpret_cast<float4*>(zs)[i] = zg[i];
    for (int i = tid; i < heads * e / 4; i += 256) reinterpret_cast<float4*>(qs)[i] = qg[i];
    __syncthreads();
    if (gamma != nullptr) {  // GroupNorm over (e / groups) channels x s positions, one warp per group
        const int cpg = e / groups, cnt = cpg * s;
        for (int g = warp; g < groups; g += 8) {
            float* base = zs + g * cpg;
            float sum = 0.f;
            for (int t = lane; t < cnt; t += 32) sum += base[(size_t)(t / cpg) * e + (t % cpg)];
            const float mean = warp_sum(sum) / cnt;
            float sq = 0.f;
            for (int t = lane; t < cnt; t += 32) {
                const float d = base[(size_t)(t / cpg) * e + (t % cpg)] - mean;
                sq += d * d;
            }
            const float rstd = rsqrtf(warp_sum(sq) / cnt + eps);
            for (int t = lane; t < cnt; t += 32) {
                const int ci = t % cpg;
                float* q = base + (size_t)(t / cpg) * e + ci;
                *q = (*q - mean) * rstd * gamma[g * cpg + ci] + beta[g * cpg + ci];
            }
        }
        __syncthreads();
    }

    // scores[h][p] = qk[h] . z[p]   (1/sqrt(d) and W_k are folded into qk by the caller)
    for (int idx = warp; idx < heads * s; idx += 8) {
        const int h = idx / s, p = idx - h * s;
        const float4* a = reinterpret_cast<const float4*>(qs + (size_t)h * e);
        const float4* b = reinterpret_cast<const float4*>(zs + (size_t)p * e);
        float acc = 0.f;
        for (int i = lane; i < e / 4; i += 32) {
            const float4 u = a[i], v = b[i];
            acc = fmaf(u.x, v.x, acc); acc = fmaf(u.y, v.y, acc);
            acc = fmaf(u.z, v.z, acc); acc = fmaf(u.w, v.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) sc[h][p] = acc;
    }
    __syncthreads();
    // softmax over positions: warp h, lane p
    if (warp < heads) {
        const float v = lane < s ? sc[warp][lane] : -INFINITY;
        const float mx = warp_max(v);
        const float ex = lane < s ? expf(v - mx) : 0.f;
        const float den = warp_sum(ex);
        // training: dropout on the attention probabilities (mask holds 0 or 1 / (1 - p) per (frame, head, position), as
        // torch's multi_head_attention_forward applies it; the returned weights are the head mean of the dropped ones)
        const float keep = mask != nullptr && lane < s ? mask[((long long)f * heads + warp) * s + lane] : 1.f;
        const float pt = lane < s ? ex / den * keep : 0.f;
        if (lane < s) sc[warp][lane] = pt;
        if (psum != nullptr) {  // sum of the kept probabilities: the weight of the value bias (1 without dropout)
            const float t = warp_sum(pt);
            if (lane == 0) psum[(long long)f * heads + warp] = t;
        }
    }
    __syncthreads();
    // head-averaged weights and the entropy term of Categorical(probs = w + 1e-6)
    if (warp == 0) {
        float w = 0.f;
        if (lane < s) {
            for (int h = 0; h < heads; ++h) w += sc[h][lane];
            w /= heads;
            wts[(long long)f * s + lane] = w;
        }
        const float pr = lane < s ? w + 1e-6f : 0.f;
        const float tot = warp_sum(pr);
        const float pn = pr / tot;
        const float term = lane < s ? -pn * logf(pn) : 0.f;
        const float ent = warp_sum(term);
        if (lane == 0 && entropy) entropy[f] = ent;
    }
    // zbar[h][c] = sum_p a[h][p] z[p][c]
    for (int c = tid; c < e; c += 256) {
        for (int h = 0; h < heads; ++h) {
            float acc = 0.f;
            for (int p = 0; p < s; ++p) acc = fmaf(sc[h][p], zs[(size_t)p * e + c], acc);
            if (zbar) zbar[((long long)f * heads + h) * e + c] = acc;
            if (zbar3) {
                const __nv_bfloat16 hi = __float2bfloat16_rn(acc);
                __nv_bfloat16* o = zbar3 + ((long long)f * heads + h) * 3 * e + c;
                o[0] = hi;
                o[e] = hi;
                o[2 * e] = __float2bfloat16_rn(acc - __bfloat162float(hi));
            }
        }
    }
}

__global__ void attn_weighted_sum_kernel(const float* __restrict__ a, const float* __restrict__ wts,
                                         float* __restrict__ y, int f, int s, int c, int y_ld) {
    const long long total = (long long)f * c;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int fi = (int)(t / c), ci = (int)(t - (long long)fi * c);
        const float* ap = a + (long long)fi * s * c + ci;
        const float* wp = wts + (long long)fi * s;
        // same order as torch: products first, then the sum over positions
        float acc = 0.f;
        for (int p = 0; p < s; ++p) acc += ap[(long long)p * c] * wp[p];
        y[(long long)fi * y_ld + ci] = acc;
    }
}

__global__ void softmax_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int cols) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long long)row * cols;
    float mx = -INFINITY;
    for (int i = lane; i < cols; i += 32) mx = fmaxf(mx, xr[i]);
    mx = warp_max(mx);
    float den = 0.f;
    for (int i = lane; i < cols; i += 32) den += expf(xr[i] - mx);
    den = warp_sum(den);
    for (int i = lane; i < cols; i += 32) y[(long long)row * cols + i] = expf(xr[i] - mx) / den;
}

__global__ void consensus_kernel(const float* __restrict__ x, float* __restrict__ y, int b, int nseg, int c) {
    const long long total = (long long)b * c;
    const float inv = 1.0f / nseg;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int bi = (int)(t / c), ci = (int)(t - (long long)bi * c);
        const float* p = x + (long long)bi * nseg * c + ci;
        float acc = 0.f;
        for (int n = 0; n < nseg; ++n) acc += p[(long long)n * c];
        y[t] = acc * inv;
    }
}

static int small_grid(long long total) {
    long long want = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

// ---------------------------------------------------------------------------------------------
// fp32 -> three bf16 column blocks [hi | hi | lo] per group of `group` columns (hi = bf16(x), lo = bf16(x - hi)).
// With weights packed [hi | lo | hi] a bf16 tensor-core GEMM over the tripled K computes
// x_hi*w_hi + x_hi*w_lo + x_lo*w_hi = x*w to ~2^-16 relative: fp32-class accuracy for the head GEMMs.
// ---------------------------------------------------------------------------------------------
__global__ void split_bf16x3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long total4,
                                    int k, int ldx, int group, long long out_ld) {
    const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= total4) return;
    const int k4 = k >> 2;
    const long long row = i4 / k4;
    const int c = (int)(i4 - row * k4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + row * ldx + c);
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        hi[j] = __float2bfloat16_rn(f[j]);
        lo[j] = __float2bfloat16_rn(f[j] - __bfloat162float(hi[j]));
    }
    const int g = c / group, j0 = c - g * group;
    __nv_bfloat16* o = out + row * out_ld + (long long)g * 3 * group + j0;
    const uint2 h2 = *reinterpret_cast<const uint2*>(hi), l2 = *reinterpret_cast<const uint2*>(lo);
    *reinterpret_cast<uint2*>(o) = h2;
    *reinterpret_cast<uint2*>(o + group) = h2;
    *reinterpret_cast<uint2*>(o + 2 * group) = l2;
}

// y[m][j] = act(sum_s partial[s][m][j] + bias[(m % period)][j]): second half of a split-K GEMM
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ bias, float* __restrict__ y,
                                     int m, int n, int ldy, int ksplit, int bias_period, int relu) {
    const int n4 = n >> 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)m * n4) return;
    const int row = (int)(i / n4), c = (int)(i - (long long)row * n4) * 4;
    float4 a = bias ? __ldg(reinterpret_cast<const float4*>(bias + (size_t)(row % bias_period) * n + c)) : make_float4(0, 0, 0, 0);
    for (int s = 0; s < ksplit; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(partial + ((size_t)s * m + row) * n + c);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    *reinterpret_cast<float4*>(y + (size_t)row * ldy + c) = a;
}

// The same reduction with the rest of the head chain folded in: the activated rows may be averaged over groups of
// `nseg` consecutive rows (segment consensus, core/models/model.py:194-199 -- the mean commutes with the affine classifier
// that follows, so the consensus moves in front of it), and the result is written as fp32 (y, optional) and / or as the
// bf16 [hi | hi | lo] operand of the next tensor-core GEMM (split, three blocks of n columns per row, optional).
__global__ void splitk_reduce_ex_kernel(const float* __restrict__ partial, const float* __restrict__ bias, float* __restrict__ y,
                                        int m, int n, int ldy, int ksplit, int bias_period, int relu, int nseg,
                                        __nv_bfloat16* __restrict__ split, long long split_ld) {
    const int n4 = n >> 2;
    const int mo = m / nseg;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)mo * n4) return;
    const int orow = (int)(i / n4), c = (int)(i - (long long)orow * n4) * 4;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int sgm = 0; sgm < nseg; ++sgm) {
        const int row = orow * nseg + sgm;
        float4 a = bias ? __ldg(reinterpret_cast<const float4*>(bias + (size_t)(row % bias_period) * n + c)) : make_float4(0, 0, 0, 0);
        for (int s = 0; s < ksplit; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(partial + ((size_t)s * m + row) * n + c);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    if (nseg > 1) {
        const float inv = 1.0f / nseg;
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    }
    if (y != nullptr) *reinterpret_cast<float4*>(y + (size_t)orow * ldy + c) = acc;
    if (split != nullptr) {
        const float f[4] = {acc.x, acc.y, acc.z, acc.w};
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = __float2bfloat16_rn(f[j]);
            lo[j] = __float2bfloat16_rn(f[j] - __bfloat162float(hi[j]));
        }
        __nv_bfloat16* o = split + (long long)orow * split_ld + c;
        const uint2 h2 = *reinterpret_cast<const uint2*>(hi), l2 = *reinterpret_cast<const uint2*>(lo);
        *reinterpret_cast<uint2*>(o) = h2;
        *reinterpret_cast<uint2*>(o + n) = h2;
        *reinterpret_cast<uint2*>(o + 2 * n) = l2;
    }
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_splitk_reduce_ex(const float* partial, const float* bias, float* y, int m, int n, int ldy, int ksplit,
                                    int bias_period, int relu, int nseg, void* split, long long split_ld, void* stream) {
    TBN_CHECK_ARG(partial && (y || split) && m > 0 && n > 0 && n % 4 == 0 && ksplit >= 1 && bias_period >= 1 && nseg >= 1 &&
                      m % nseg == 0,
                  "splitk_reduce_ex: bad arguments");
    TBN_CHECK_ARG(y == nullptr || ldy % 4 == 0, "splitk_reduce_ex: ldy %% 4");
    TBN_CHECK_ARG(split == nullptr || (split_ld >= 3LL * n && split_ld % 4 == 0 && ((uintptr_t)split & 7) == 0),
                  "splitk_reduce_ex: bad split operand");
    const long long total = (long long)(m / nseg) * (n / 4);
    splitk_reduce_ex_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(partial, bias, y, m, n, ldy, ksplit, bias_period, relu,
                                                                              nseg, reinterpret_cast<__nv_bfloat16*>(split), split_ld);
    TBN_LAUNCH_CHECK("splitk_reduce_ex launch");
    return 0;
}

extern "C" int tbn_splitk_reduce(const float* partial, const float* bias, float* y, int m, int n, int ldy, int ksplit,
                                 int bias_period, int relu, void* stream) {
    TBN_CHECK_ARG(partial && y && m > 0 && n > 0 && n % 4 == 0 && ldy % 4 == 0 && ksplit >= 1 && bias_period >= 1,
                  "splitk_reduce: bad arguments");
    const long long total = (long long)m * (n / 4);
    splitk_reduce_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(partial, bias, y, m, n, ldy, ksplit, bias_period, relu);
    TBN_LAUNCH_CHECK("splitk_reduce launch");
    return 0;
}

extern "C" int tbn_split_bf16x3_ld(const float* x, void* out, long long out_ld, int m, int k, int ldx, int group, void* stream) {
    TBN_CHECK_ARG(x && out && m > 0 && k > 0, "split_bf16x3: bad arguments");
    TBN_CHECK_ARG(k % 4 == 0 && ldx % 4 == 0 && group % 4 == 0 && k % group == 0, "split_bf16x3: k, ldx, group %% 4; k %% group");
    TBN_CHECK_ARG(out_ld >= 3LL * k && out_ld % 4 == 0, "split_bf16x3: output leading dimension must be >= 3k and %% 4");
    TBN_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 7) == 0, "split_bf16x3: alignment");
    const long long total4 = (long long)m * (k / 4);
    split_bf16x3_kernel<<<cdiv(total4, 256), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(out), total4, k,
                                                                           ldx, group, out_ld);
    TBN_LAUNCH_CHECK("split_bf16x3 launch");
    return 0;
}

extern "C" int tbn_split_bf16x3(const float* x, void* out, int m, int k, int ldx, int group, void* stream) {
    return tbn_split_bf16x3_ld(x, out, 3LL * k, m, k, ldx, group, stream);
}

extern "C" int tbn_linear(const float* x, const float* wt, const float* bias, float* y, int m, int n, int k, int ldx,
                          int ldw, int ldy, int bias_period, int relu, void* stream) {
    LinearArgs a{x, wt, bias, y, m, n, k, ldx, ldw, ldy, bias_period, relu, 0, 0, 0, 0};
    return launch_linear(a, 1, as_stream(stream));
}

extern "C" int tbn_linear_grouped(const float* x, const float* wt, const float* bias, float* y, int m, int n, int k,
                                  int ldx, int ldw, int ldy, int groups, long long gx, long long gw, long long gb,
                                  long long gy, int relu, void* stream) {
    TBN_CHECK_ARG(groups >= 1 && groups <= 65535, "linear_grouped: bad group count %d", groups);
    LinearArgs a{x, wt, bias, y, m, n, k, ldx, ldw, ldy, 1, relu, gx, gw, gb, gy};
    return launch_linear(a, groups, as_stream(stream));
}

extern "C" int tbn_groupnorm(float* z, const float* gamma, const float* beta, int f, int s, int c, int groups,
                             float eps, void* stream) {
    TBN_CHECK_ARG(z && gamma && beta && f > 0 && s > 0 && groups > 0 && c % groups == 0, "groupnorm: bad arguments");
    const long long warps = (long long)f * groups;
    groupnorm_kernel<<<cdiv(warps * 32, 256), 256, 0, as_stream(stream)>>>(z, gamma, beta, f, s, c, groups, eps);
    TBN_LAUNCH_CHECK("groupnorm launch");
    return 0;
}

static int launch_mha_core(const float* qk, const float* z, float* zbar, float* wts, float* entropy, int f, int s, int e, int heads,
                           const float* gamma, const float* beta, int groups, float eps, void* zbar3, void* stream,
                           const float* mask = nullptr, float* psum = nullptr);

extern "C" int tbn_mha_core(const float* qk, const float* z, float* zbar, float* wts, float* entropy, int f, int s,
                            int e, int heads, void* stream) {
    TBN_CHECK_ARG(zbar != nullptr, "mha_core: null pointer");
    return launch_mha_core(qk, z, zbar, wts, entropy, f, s, e, heads, nullptr, nullptr, 1, 0.f, nullptr, stream);
}

extern "C" int tbn_mha_core_gn(const float* qk, const float* z, const float* gamma, const float* beta, int groups, float eps,
                               float* zbar, void* zbar3, float* wts, float* entropy, int f, int s, int e, int heads,
                               void* stream) {
    TBN_CHECK_ARG(gamma && beta && groups > 0 && e % groups == 0, "mha_core_gn: bad GroupNorm arguments");
    TBN_CHECK_ARG(zbar != nullptr || zbar3 != nullptr, "mha_core_gn: no output for the weighted value sums");
    return launch_mha_core(qk, z, zbar, wts, entropy, f, s, e, heads, gamma, beta, groups, eps, zbar3, stream);
}

extern "C" int tbn_mha_core_train(const float* qk, const float* z, const float* mask, float* zbar, float* wts, float* entropy,
                                  float* psum, int f, int s, int e, int heads, void* stream) {
    TBN_CHECK_ARG(zbar != nullptr, "mha_core_train: null pointer");
    return launch_mha_core(qk, z, zbar, wts, entropy, f, s, e, heads, nullptr, nullptr, 1, 0.f, nullptr, stream, mask, psum);
}

static int launch_mha_core(const float* qk, const float* z, float* zbar, float* wts, float* entropy, int f, int s, int e, int heads,
                           const float* gamma, const float* beta, int groups, float eps, void* zbar3, void* stream,
                           const float* mask, float* psum) {
    TBN_CHECK_ARG(qk && z && wts && f > 0, "mha_core: null pointer");
    TBN_CHECK_ARG(s >= 1 && s <= MHA_MAX_S && heads >= 1 && heads <= MHA_MAX_H && e % 4 == 0,
                  "mha_core: need 1<=S<=32, 1<=heads<=8, E%%4==0 (got S=%d heads=%d E=%d)", s, heads, e);
    const size_t smem = (size_t)(s + heads) * e * sizeof(float);
    TBN_CHECK_ARG(smem <= 200 * 1024, "mha_core: S*E too large for shared memory");
    static DeviceOnce once;
    int once_dev = 0;
    if (smem > 48 * 1024 && once.needed(&once_dev)) {
        cudaError_t er = cudaFuncSetAttribute(mha_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (er != cudaSuccess) return cuda_fail(er, "mha_core smem opt-in");
        once.done(once_dev);
    }
    mha_core_kernel<<<f, 256, smem, as_stream(stream)>>>(qk, z, zbar, wts, entropy, s, e, heads, gamma, beta, groups, eps,
                                                         reinterpret_cast<__nv_bfloat16*>(zbar3), mask, psum);
    TBN_LAUNCH_CHECK("mha_core launch");
    return 0;
}

// rows [0, rows) of an (k*rows, cols) matrix with leading dimension ld are copied to rows [j*rows, (j+1)*rows), j = 1..k-1
__global__ void repeat_rows_kernel(float* __restrict__ x, int rows, int cols4, long long ld, int k) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols4) return;
    const int r = (int)(i / cols4), c = (int)(i - (long long)r * cols4);
    const float4 v = *reinterpret_cast<const float4*>(x + (long long)r * ld + 4 * c);
    for (int j = 1; j < k; ++j) *reinterpret_cast<float4*>(x + ((long long)j * rows + r) * ld + 4 * c) = v;
}

extern "C" int tbn_repeat_rows(float* x, int rows, int cols, int ld, int k, void* stream) {
    TBN_CHECK_ARG(x && rows > 0 && cols > 0 && k >= 1, "repeat_rows: bad arguments");
    TBN_CHECK_ARG(cols % 4 == 0 && ld % 4 == 0 && ((uintptr_t)x & 15) == 0, "repeat_rows: cols, ld %% 4 and 16-byte alignment required");
    if (k == 1) return 0;
    const long long total = (long long)rows * (cols / 4);
    repeat_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(x, rows, cols / 4, ld, k);
    TBN_LAUNCH_CHECK("repeat_rows launch");
    return 0;
}

extern "C" int tbn_attn_weighted_sum(const float* a, const float* wts, float* y, int f, int s, int c, int y_ld,
                                     void* stream) {
    TBN_CHECK_ARG(a && wts && y && f > 0 && s > 0 && c > 0 && y_ld >= c, "attn_weighted_sum: bad arguments");
    attn_weighted_sum_kernel<<<small_grid((long long)f * c), 256, 0, as_stream(stream)>>>(a, wts, y, f, s, c, y_ld);
    TBN_LAUNCH_CHECK("attn_weighted_sum launch");
    return 0;
}

extern "C" int tbn_softmax_rows(const float* x, float* y, int rows, int cols, void* stream) {
    TBN_CHECK_ARG(x && y && rows > 0 && cols > 0, "softmax_rows: bad arguments");
    softmax_rows_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, as_stream(stream)>>>(x, y, rows, cols);
    TBN_LAUNCH_CHECK("softmax_rows launch");
    return 0;
}

extern "C" int tbn_consensus(const float* x, float* y, int b, int nseg, int c, void* stream) {
    TBN_CHECK_ARG(x && y && b > 0 && nseg > 0 && c > 0, "consensus: bad arguments");
    consensus_kernel<<<small_grid((long long)b * c), 256, 0, as_stream(stream)>>>(x, y, b, nseg, c);
    TBN_LAUNCH_CHECK("consensus launch");
    return 0;
}
