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
