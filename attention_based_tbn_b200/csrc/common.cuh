// Shared helpers of libtbn_sm100.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/tbn_sm100.h"

namespace tbn {

// ---- error plumbing: nothing throws across the C ABI ------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define TBN_CHECK_ARG(cond, ...)          \
    do {                                  \
        if (!(cond)) {                    \
            tbn::set_error(__VA_ARGS__);  \
            return 1;                     \
        }                                 \
    } while (0)

#define TBN_LAUNCH_CHECK(what)                                   \
    do {                                                         \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) return tbn::cuda_fail(e__, what); \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- storage-type helpers ---------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive channels as one vector (32 B for fp32, 16 B for bf16).
template <typename T>
struct Vec8;
template <>
struct Vec8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = *reinterpret_cast<const float4*>(p);
        b = *reinterpret_cast<const float4*>(p + 4);
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = a;
        *reinterpret_cast<float4*>(p + 4) = b;
    }
    __device__ __forceinline__ void get(float* v) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ __forceinline__ void set(const float* v) {
        a = make_float4(v[0], v[1], v[2], v[3]);
        b = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <>
struct Vec8<__nv_bfloat16> {
    uint4 u;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = u; }
    __device__ __forceinline__ void get(float* v) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    __device__ __forceinline__ void set(const float* v) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

int sm_count();

// One-time per-DEVICE configuration (cudaFuncSetAttribute applies to the current device only, so a process or thread
// that moves to a second GPU must opt in again there): bit d of the mask = device d is configured.
struct DeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool needed(int* dev) {
        cudaGetDevice(dev);
        return ((mask.load(std::memory_order_acquire) >> (*dev & 63)) & 1ull) == 0;
    }
    void done(int dev) { mask.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

}  // namespace tbn
