// Hardware experiment (sm_100a): what does a TMA tensor load cost -- per instruction, per 128-byte row, or per byte?
// One elected thread per CTA streams boxes of several shapes from an L2-resident tensor into a shared-memory ring
// (it only waits for a box's bytes when its ring slot comes round again), 148 CTAs, and reports bytes / clock / SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_rate tma_rate.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spins = 0; !ok; ++spins) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 24)) { printf("timeout\n"); __trap(); }
    }
}
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// tensor viewed as {64 ch (inner, 128 B), rows, slices}: row pitch `pitch` bytes, slice stride 128 B
__global__ void __launch_bounds__(128) kern(const __grid_constant__ CUtensorMap tm, int box_bytes, int rows_per_box, int nslice,
                                            int stages, int iters, int total_rows, unsigned long long* cycles) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bars = base + (uint32_t)stages * (uint32_t)box_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(bars + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        int row = (blockIdx.x * 977) % (total_rows - rows_per_box);
        for (int i = 0; i < iters; ++i) {
            const int s = i % stages;
            if (i >= stages) mbar_wait(bars + 8u * s, ((i / stages) - 1) & 1);  // previous box in this slot has landed
            mbar_expect_tx(bars + 8u * s, (uint32_t)box_bytes);
            tma3(base + (uint32_t)s * (uint32_t)box_bytes, &tm, bars + 8u * s, 0, row, 0);
            row += rows_per_box;
            if (row > total_rows - rows_per_box) row = 0;
        }
        for (int i = iters; i < iters + stages; ++i) {  // drain
            const int s = i % stages;
            if (i >= stages) mbar_wait(bars + 8u * s, ((i / stages) - 1) & 1);
        }
        cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fp;
    const int total_rows = 65536;           // x 256 channels bf16 = 32 MB: L2 resident
    const int chan = 256;
    __nv_bfloat16* d; CK(cudaMalloc(&d, (size_t)total_rows * chan * 2)); CK(cudaMemset(d, 0, (size_t)total_rows * chan * 2));
    unsigned long long* dc; CK(cudaMalloc(&dc, 148 * 8));
    int sms = 148;
    struct Cfg { int rows, nslice; const char* name; } cfgs[] = {
        {32, 1, "32 rows x 128 B   (4 KB)"}, {64, 1, "64 rows x 128 B   (8 KB)"}, {128, 1, "128 rows x 128 B (16 KB)"},
        {256, 1, "256 rows x 128 B (32 KB)"}, {128, 2, "2 x 128 rows x 128 B (32 KB, one instr)"}, {64, 2, "2 x 64 rows (16 KB)"},
        {128, 3, "3 x 128 rows (48 KB)"},
    };
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    for (auto& c : cfgs) {
        CUtensorMap tm;
        cuuint64_t dims[3] = {64, (cuuint64_t)total_rows, (cuuint64_t)(chan / 64)};
        cuuint64_t str[2] = {(cuuint64_t)chan * 2, 128};
        cuuint32_t box[3] = {64, (cuuint32_t)c.rows, (cuuint32_t)c.nslice};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode failed %d\n", r); return 2; }
        const int box_bytes = c.rows * c.nslice * 128;
        for (int stages : {2, 4, 8}) {
            if ((size_t)stages * box_bytes > 200 * 1024) continue;
            const int iters = 400;
            const size_t smem = (size_t)stages * box_bytes + 1024 + 8 * stages + 64;
            kern<<<sms, 128, smem>>>(tm, box_bytes, c.rows, c.nslice, stages, iters, total_rows, dc);  // warm L2
            CK(cudaDeviceSynchronize());
            kern<<<sms, 128, smem>>>(tm, box_bytes, c.rows, c.nslice, stages, iters, total_rows, dc);
            CK(cudaDeviceSynchronize());
            unsigned long long hc[148]; CK(cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost));
            double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)hc[i]; avg /= sms;
            printf("%-42s stages %d: %7.0f cycles/instr  %6.1f B/clk/SM\n", c.name, stages, avg / iters, (double)box_bytes * iters / avg);
        }
    }
    return 0;
}
