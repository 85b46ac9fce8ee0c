// Audio front-end on the GPU: log-power STFT of the audio windows that feed the Audio backbone
// (reference core/dataset/dataset.py:459-497, Video_Dataset._get_spectrogram: librosa.stft(n_fft=511, hop=120,
// win_length=240, window="hann", center=True, pad_mode="constant") then log(|X|^2 + 1e-6)).
//
// Only `win` of the n_fft frame samples are non-zero after the centred Hann window, so the transform is the real GEMM
//     [frames x win] * [win x 2*bins]      (window * cos | -window * sin, built once on the host in float64)
// and runs on the tensor cores through the same split-operand trick as the head GEMMs: samples go in as bf16
// [hi | hi | lo], the DFT matrix as [hi | lo | hi] (fp32-class accuracy, ~2^-16 per product).
//   tbn_stft_frames    gathers the frames (centre padding = zeros outside the window segment) and splits them,
//   tbn_conv_tc (k=1)  is the GEMM (include/tbn_sm100.h),
//   tbn_stft_logpower  squares, adds, takes the logarithm and transposes to the (n, 1, bins, frames) NCHW spectrogram.
#include "common.cuh"

namespace tbn {

// One thread = 4 consecutive window samples of one frame.  a3 row layout: [hi (kp) | hi (kp) | lo (kp)], kp = win padded
// to 64; columns >= win are zero.
__global__ void stft_frames_kernel(const float* __restrict__ samples, long long row_stride, int seg_len, int frames, int hop,
                                   int win, int shift, int kp, long long total, __nv_bfloat16* __restrict__ a3) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int q = kp >> 2;
    const long long row = i / q;
    const int j0 = (int)(i - row * q) * 4;
    const int s = (int)(row / frames), t = (int)(row - (long long)s * frames);
    const float* src = samples + (long long)s * row_stride;
    const int local0 = t * hop - shift + j0;  // sample index inside the segment; outside [0, seg_len) = centre padding
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int l = local0 + j;
        const float v = (j0 + j < win && l >= 0 && l < seg_len) ? __ldg(src + l) : 0.f;
        hi[j] = __float2bfloat16_rn(v);
        lo[j] = __float2bfloat16_rn(v - __bfloat162float(hi[j]));
    }
    __nv_bfloat16* o = a3 + row * 3LL * kp + j0;
    const uint2 h2 = *reinterpret_cast<const uint2*>(hi), l2 = *reinterpret_cast<const uint2*>(lo);
    *reinterpret_cast<uint2*>(o) = h2;
    *reinterpret_cast<uint2*>(o + kp) = h2;
    *reinterpret_cast<uint2*>(o + 2 * kp) = l2;
}

// g: (n*frames, 2*bins) fp32, columns [0, bins) = Re, [bins, 2*bins) = Im.  out: (n, bins, frames) fp32.
// 32 x 32 tile transposed through shared memory: reads coalesced along the bins, writes along the frames.
__global__ void stft_logpower_kernel(const float* __restrict__ g, float* __restrict__ out, int frames, int bins, float eps) {
    __shared__ float tile[32][33];
    const int s = blockIdx.z;
    const int t0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int t = t0 + r, k = k0 + tx;
        float v = 0.f;
        if (t < frames && k < bins) {
            const float* row = g + ((long long)s * frames + t) * (2LL * bins);
            const float re = row[k], im = row[bins + k];
            v = logf(re * re + im * im + eps);
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, t = t0 + tx;
        if (k < bins && t < frames) out[((long long)s * bins + k) * frames + t] = tile[tx][r];
    }
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_stft_frames(const float* samples, long long row_stride, int n, int seg_len, int n_fft, int win, int hop,
                               void* a3, void* stream) {
    TBN_CHECK_ARG(samples && a3 && n > 0 && seg_len > 0, "stft_frames: bad arguments");
    TBN_CHECK_ARG(n_fft > 0 && win > 0 && win <= n_fft && hop > 0, "stft_frames: need 0 < win <= n_fft, hop > 0");
    TBN_CHECK_ARG(((uintptr_t)a3 & 7) == 0, "stft_frames: output must be 8-byte aligned");
    const int pad = n_fft / 2;                                   // librosa center=True pads n_fft // 2 on both sides
    const int frames = 1 + (seg_len + 2 * pad - n_fft) / hop;    // librosa.util.frame
    TBN_CHECK_ARG(seg_len + 2 * pad >= n_fft && frames > 0, "stft_frames: segment shorter than one frame");
    const int shift = pad - (n_fft - win) / 2;                   // the window sits centred in the n_fft frame (pad_center)
    const int kp = (win + 63) / 64 * 64;
    const long long total = (long long)n * frames * (kp / 4);
    TBN_CHECK_ARG((long long)n * frames < (1LL << 31), "stft_frames: too many frames");
    stft_frames_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(samples, row_stride, seg_len, frames, hop, win, shift, kp,
                                                                         total, reinterpret_cast<__nv_bfloat16*>(a3));
    TBN_LAUNCH_CHECK("stft_frames launch");
    return 0;
}

extern "C" int tbn_stft_logpower(const float* g, float* out, int n, int frames, int bins, float eps, void* stream) {
    TBN_CHECK_ARG(g && out && n > 0 && frames > 0 && bins > 0, "stft_logpower: bad arguments");
    TBN_CHECK_ARG(n <= 65535, "stft_logpower: at most 65535 spectrograms per call");
    dim3 grid(cdiv(frames, 32), cdiv(bins, 32), n), block(32, 8);
    stft_logpower_kernel<<<grid, block, 0, as_stream(stream)>>>(g, out, frames, bins, eps);
    TBN_LAUNCH_CHECK("stft_logpower launch");
    return 0;
}
