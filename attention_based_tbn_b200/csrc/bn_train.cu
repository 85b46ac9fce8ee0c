// Train-mode BatchNorm2d for the training forward (reference core/tools/train.py:68-76: model.train(), so every
// BatchNorm of the backbones normalises with BATCH statistics and updates its running statistics -- the eval path's
// folded BatchNorm does not apply).  Works in place on a channel slice of an NHWC activation (fp32, the CUDA-core
// training path): pass 1 accumulates per-channel sum / sum of squares in double, pass 2 normalises, applies the affine
// transform and ReLU, and (block 0) updates running_mean / running_var with the UNBIASED variance, as torch does.
#include "common.cuh"

namespace tbn {

constexpr int BN_ROWS_PER_BLOCK = 256;
constexpr int BN_MAX_C = 1056;  // widest conv of the backbones is 352; the widest concat 1056

// sum the per-thread partials of the row lanes of a block (threads ch, ch + cpt, ...) and issue ONE atomic per channel
__device__ __forceinline__ void bn_block_reduce_add(double s0, double s1, double* sums, int c, int cc, int cpt, int lanes, int lane,
                                                    int ch, bool active) {
    __shared__ double red[2][256];
    red[0][threadIdx.x] = active ? s0 : 0.0;
    red[1][threadIdx.x] = active ? s1 : 0.0;
    __syncthreads();
    if (lane == 0 && active) {
        double a = 0.0, b = 0.0;
        for (int l = 0; l < lanes; ++l) {
            a += red[0][l * cpt + ch];
            b += red[1][l * cpt + ch];
        }
        atomicAdd(&sums[cc], a);
        atomicAdd(&sums[c + cc], b);
    }
}

__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, long long m, int c,
                                                       int cs, int co) {
    // block = 256 threads = (256 / cpt) row lanes x cpt channel threads; grid.y tiles the channels in groups of 64
    const int cpt = c < 64 ? c : 64;          // channel threads
    const int lanes = 256 / cpt;              // rows in flight
    const int ch = threadIdx.x % cpt, lane = threadIdx.x / cpt;
    const int cbase = blockIdx.y * 64;
    const bool active = lane < lanes && cbase + ch < c;
    const long long r0 = (long long)blockIdx.x * BN_ROWS_PER_BLOCK * lanes;
    double s = 0.0, ss = 0.0;
    if (active) {
        // 8 independent loads in flight per thread (rows past the end read row m - 1 and count as zero)
        for (int i = 0; i < BN_ROWS_PER_BLOCK; i += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long r = r0 + (long long)(i + j) * lanes + lane;
                v[j] = x[(r < m ? r : m - 1) * cs + co + cbase + ch];
                if (r >= m) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double d = (double)v[j];
                s += d;
                ss += d * d;
            }
            if (r0 + (long long)(i + 8) * lanes >= m) break;
        }
    }
    bn_block_reduce_add(s, ss, sums, c, cbase + ch, cpt, lanes, lane, ch, active);
}

__global__ void bn_apply_kernel(float* __restrict__ x, const double* __restrict__ sums, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                float momentum, float eps, long long m, int c, int cs, int co, int relu) {
    const long long total = m * c;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(t % c);
        const long long r = t / c;
        const double mean = sums[ch] / (double)m;
        const double var = fmax(sums[c + ch] / (double)m - mean * mean, 0.0);
        const float inv = (float)(1.0 / sqrt(var + (double)eps));
        float* p = x + r * cs + co + ch;
        float v = (*p - (float)mean) * inv * gamma[ch] + beta[ch];
        if (relu) v = fmaxf(v, 0.f);
        *p = v;
    }
    if (blockIdx.x == 0 && rmean != nullptr) {
        __syncthreads();
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            const double mean = sums[ch] / (double)m;
            const double var = fmax(sums[c + ch] / (double)m - mean * mean, 0.0);
            const double unbiased = m > 1 ? var * (double)m / (double)(m - 1) : var;
            rmean[ch] = (1.f - momentum) * rmean[ch] + momentum * (float)mean;
            rvar[ch] = (1.f - momentum) * rvar[ch] + momentum * (float)unbiased;
        }
    }
}

// Out-of-place variant for the trainable backbones: u (m, c) contiguous pre-norm conv output is KEPT (the backward needs
// it), y = relu(bn(u)) goes into a channel slice, the batch mean / rstd are saved.  Same arithmetic as bn_apply_kernel.
__global__ void bn_apply_save_kernel(const float* __restrict__ u, float* __restrict__ y, const double* __restrict__ sums,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ rmean,
                                     float* __restrict__ rvar, float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                     float momentum, float eps, long long m, int c, int cs, int co) {
    // per-channel constants once per block (the double-precision mean / variance arithmetic is not repeated per element)
    __shared__ float s_mean[BN_MAX_C], s_inv[BN_MAX_C];
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        const double mean = sums[ch] / (double)m;
        const double var = fmax(sums[c + ch] / (double)m - mean * mean, 0.0);
        s_mean[ch] = (float)mean;
        s_inv[ch] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    // (c, cs, co are multiples of 4 and m * c / 4 < 2^32 for every layer of the backbones: 4 channels per thread)
    const unsigned c4 = (unsigned)c >> 2;
    const unsigned long long total = (unsigned long long)m * c4;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned t32 = (unsigned)t;
        const unsigned r = t32 / c4, ch = (t32 - r * c4) * 4;
        const float4 uv = *reinterpret_cast<const float4*>(u + (size_t)t32 * 4);
        const float in[4] = {uv.x, uv.y, uv.z, uv.w};
        float out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = fmaxf((in[j] - s_mean[ch + j]) * s_inv[ch + j] * gamma[ch + j] + beta[ch + j], 0.f);
        *reinterpret_cast<float4*>(y + (size_t)r * cs + co + ch) = make_float4(out[0], out[1], out[2], out[3]);
    }
    if (blockIdx.x == 0) {
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            const double mean = sums[ch] / (double)m;
            const double var = fmax(sums[c + ch] / (double)m - mean * mean, 0.0);
            const double unbiased = m > 1 ? var * (double)m / (double)(m - 1) : var;
            save_mean[ch] = (float)mean;
            save_rstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
            if (rmean != nullptr) {
                rmean[ch] = (1.f - momentum) * rmean[ch] + momentum * (float)mean;
                rvar[ch] = (1.f - momentum) * rvar[ch] + momentum * (float)unbiased;
            }
        }
    }
}

// BatchNorm(batch statistics) + ReLU backward.  g = dy * [y > 0];  sums[ch] = sum g, sums[c + ch] = sum g * xhat.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ u, const float* __restrict__ dy,
                                                            const float* __restrict__ y, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, double* __restrict__ sums, long long m,
                                                            int c, int cs, int co, int dcs, int dco) {
    const int cpt = c < 64 ? c : 64;
    const int lanes = 256 / cpt;
    const int ch = threadIdx.x % cpt, lane = threadIdx.x / cpt;
    const int cbase = blockIdx.y * 64;
    const bool active = lane < lanes && cbase + ch < c;
    const int cc = cbase + ch;
    double s = 0.0, sx = 0.0;
    if (active) {
        const float mu = mean[cc], rs = rstd[cc];
        const long long r0 = (long long)blockIdx.x * BN_ROWS_PER_BLOCK * lanes;
        for (int i = 0; i < BN_ROWS_PER_BLOCK; i += 4) {
            float yv[4], gv[4], uv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // 12 independent loads in flight per thread
                const long long r = r0 + (long long)(i + j) * lanes + lane;
                const long long rr = r < m ? r : m - 1;
                yv[j] = y[rr * cs + co + cc];
                gv[j] = dy[rr * dcs + dco + cc];
                uv[j] = u[rr * c + cc];
                if (r >= m) yv[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (yv[j] > 0.f) {
                    const double g = (double)gv[j];
                    s += g;
                    sx += g * (double)((uv[j] - mu) * rs);
                }
            }
            if (r0 + (long long)(i + 4) * lanes >= m) break;
        }
    }
    bn_block_reduce_add(s, sx, sums, c, cc, cpt, lanes, lane, ch, active);
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ u, const float* __restrict__ dy, const float* __restrict__ y,
                                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const double* __restrict__ sums, float* __restrict__ du, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, long long m, int c, int cs, int co, int dcs, int dco) {
    __shared__ float s_mg[BN_MAX_C], s_mgx[BN_MAX_C], s_gr[BN_MAX_C];
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        s_mg[ch] = (float)(sums[ch] / (double)m);
        s_mgx[ch] = (float)(sums[c + ch] / (double)m);
        s_gr[ch] = gamma[ch] * rstd[ch];
    }
    __syncthreads();
    const unsigned c4 = (unsigned)c >> 2;
    const unsigned long long total = (unsigned long long)m * c4;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned t32 = (unsigned)t;
        const unsigned r = t32 / c4, ch = (t32 - r * c4) * 4;
        const float4 uv = *reinterpret_cast<const float4*>(u + (size_t)t32 * 4);
        const float4 yv = *reinterpret_cast<const float4*>(y + (size_t)r * cs + co + ch);
        const float4 dv = *reinterpret_cast<const float4*>(dy + (size_t)r * dcs + dco + ch);
        const float ui[4] = {uv.x, uv.y, uv.z, uv.w}, yi[4] = {yv.x, yv.y, yv.z, yv.w}, di[4] = {dv.x, dv.y, dv.z, dv.w};
        float out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float g = yi[j] > 0.f ? di[j] : 0.f;
            const float xh = (ui[j] - mean[ch + j]) * rstd[ch + j];
            out[j] = s_gr[ch + j] * (g - s_mg[ch + j] - xh * s_mgx[ch + j]);
        }
        *reinterpret_cast<float4*>(du + (size_t)t32 * 4) = make_float4(out[0], out[1], out[2], out[3]);
    }
    if (blockIdx.x == 0) {
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            if (dbeta) dbeta[ch] = (float)sums[ch];
            if (dgamma) dgamma[ch] = (float)sums[c + ch];
        }
    }
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_bn_train_fwd(const float* u, float* y, int y_cs, int y_co, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, float* save_mean, float* save_rstd, void* workspace,
                                float momentum, float eps, long long m, int c, void* stream) {
    TBN_CHECK_ARG(u && y && gamma && beta && save_mean && save_rstd && workspace && m > 0 && c > 0 && y_co >= 0 && y_co + c <= y_cs,
                  "bn_train_fwd: bad arguments");
    TBN_CHECK_ARG(((uintptr_t)workspace & 7) == 0, "bn_train_fwd: workspace must hold 2*c doubles (8-byte aligned)");
    cudaStream_t st = as_stream(stream);
    double* sums = reinterpret_cast<double*>(workspace);
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * c, st);
    if (e != cudaSuccess) return cuda_fail(e, "bn_train_fwd memset");
    const int cpt = c < 64 ? c : 64;
    const int lanes = 256 / cpt;
    const long long rows_per_block = (long long)BN_ROWS_PER_BLOCK * lanes;
    dim3 grid((unsigned)((m + rows_per_block - 1) / rows_per_block), (unsigned)((c + 63) / 64));
    bn_stats_kernel<<<grid, 256, 0, st>>>(u, sums, m, c, c, 0);
    TBN_LAUNCH_CHECK("bn_stats launch");
    TBN_CHECK_ARG(c % 4 == 0 && c <= BN_MAX_C && y_cs % 4 == 0 && y_co % 4 == 0 && m * (c / 4) < (1ll << 32),
                  "bn_train_fwd: c (<= 1056), y_cs, y_co must be multiples of 4");
    long long blocks = (m * (c / 4) + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    bn_apply_save_kernel<<<(unsigned)blocks, 256, 0, st>>>(u, y, sums, gamma, beta, running_mean, running_var, save_mean, save_rstd,
                                                           momentum, eps, m, c, y_cs, y_co);
    TBN_LAUNCH_CHECK("bn_apply_save launch");
    return 0;
}

extern "C" int tbn_bn_train_bwd(const float* u, const float* dy, int dy_cs, int dy_co, const float* y, int y_cs, int y_co,
                                const float* gamma, const float* save_mean, const float* save_rstd, float* du, float* dgamma,
                                float* dbeta, void* workspace, long long m, int c, void* stream) {
    TBN_CHECK_ARG(u && dy && y && gamma && save_mean && save_rstd && du && workspace && m > 0 && c > 0, "bn_train_bwd: bad arguments");
    TBN_CHECK_ARG(dy_co + c <= dy_cs && y_co + c <= y_cs, "bn_train_bwd: channel slice outside the buffer");
    cudaStream_t st = as_stream(stream);
    double* sums = reinterpret_cast<double*>(workspace);
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * c, st);
    if (e != cudaSuccess) return cuda_fail(e, "bn_train_bwd memset");
    const int cpt = c < 64 ? c : 64;
    const int lanes = 256 / cpt;
    const long long rows_per_block = (long long)BN_ROWS_PER_BLOCK * lanes;
    dim3 grid((unsigned)((m + rows_per_block - 1) / rows_per_block), (unsigned)((c + 63) / 64));
    bn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(u, dy, y, save_mean, save_rstd, sums, m, c, y_cs, y_co, dy_cs, dy_co);
    TBN_LAUNCH_CHECK("bn_bwd_reduce launch");
    TBN_CHECK_ARG(c % 4 == 0 && c <= BN_MAX_C && y_cs % 4 == 0 && y_co % 4 == 0 && dy_cs % 4 == 0 && dy_co % 4 == 0 && m * (c / 4) < (1ll << 32),
                  "bn_train_bwd: c and the channel strides / offsets must be multiples of 4");
    long long blocks = (m * (c / 4) + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    bn_bwd_apply_kernel<<<(unsigned)blocks, 256, 0, st>>>(u, dy, y, gamma, save_mean, save_rstd, sums, du, dgamma, dbeta, m, c, y_cs, y_co,
                                                          dy_cs, dy_co);
    TBN_LAUNCH_CHECK("bn_bwd_apply launch");
    return 0;
}

extern "C" int tbn_bn_train_relu(float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                 void* workspace, float momentum, float eps, long long m, int c, int cs, int co, int relu,
                                 void* stream) {
    TBN_CHECK_ARG(x && gamma && beta && workspace && m > 0 && c > 0 && co >= 0 && co + c <= cs, "bn_train_relu: bad arguments");
    TBN_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_train_relu: running_mean / running_var go together");
    TBN_CHECK_ARG(((uintptr_t)workspace & 7) == 0, "bn_train_relu: workspace must hold 2*c doubles (8-byte aligned)");
    cudaStream_t st = as_stream(stream);
    double* sums = reinterpret_cast<double*>(workspace);
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * c, st);
    if (e != cudaSuccess) return cuda_fail(e, "bn_train_relu memset");
    const int cpt = c < 64 ? c : 64;
    const int lanes = 256 / cpt;
    const long long rows_per_block = (long long)BN_ROWS_PER_BLOCK * lanes;
    dim3 grid((unsigned)((m + rows_per_block - 1) / rows_per_block), (unsigned)((c + 63) / 64));
    bn_stats_kernel<<<grid, 256, 0, st>>>(x, sums, m, c, cs, co);
    TBN_LAUNCH_CHECK("bn_stats launch");
    long long blocks = (m * c + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    bn_apply_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, sums, gamma, beta, running_mean, running_var, momentum, eps, m, c, cs, co, relu);
    TBN_LAUNCH_CHECK("bn_apply launch");
    return 0;
}
