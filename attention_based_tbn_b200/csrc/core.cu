// Error plumbing, version and device queries of libtbn_sm100.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace tbn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return 100 + (int)e;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace tbn

// Diagnostic: one thread writes %globaltimer (ns) into *slot.  Launched around other kernels of a stream / captured
// graph it gives a device-side timeline (scripts/timeline_step.py).
__global__ void stamp_kernel(unsigned long long* slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}

extern "C" int tbn_debug_stamp(void* slot, void* stream) {
    TBN_CHECK_ARG(slot != nullptr, "debug_stamp: null slot");
    stamp_kernel<<<1, 1, 0, tbn::as_stream(stream)>>>(reinterpret_cast<unsigned long long*>(slot));
    TBN_LAUNCH_CHECK("debug_stamp launch");
    return 0;
}

extern "C" int tbn_version(void) { return 100; }

extern "C" const char* tbn_last_error(void) { return tbn::g_err; }

extern "C" int tbn_device_info(int* sms, int* major, int* minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return tbn::cuda_fail(e, "cudaGetDevice");
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return tbn::cuda_fail(e, "cudaGetDeviceProperties");
    if (sms) *sms = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    return 0;
}
