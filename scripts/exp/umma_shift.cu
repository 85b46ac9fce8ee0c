// Hardware experiment (sm_100a): does a tcgen05.mma K-major SWIZZLE_128B shared-memory descriptor whose start
// address is NOT 1024-byte aligned (shifted by whole 128-byte rows) and whose stride-byte-offset is not 1024
// read the rows one expects?  This decides whether a 3x3 convolution can feed all nine filter taps from ONE
// halo tile in shared memory (tap (r,s) = the same tile shifted by r*pitch+s rows).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_shift umma_shift.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Cfg { int shift, sbo, bo_mode; };
constexpr int NCFG_MAX = 64;
struct Cfgs { int n; Cfg c[NCFG_MAX]; };

constexpr int N = 64;   // MMA N
constexpr int ROWS = 256;

__global__ void __launch_bounds__(128) kern(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Cfgs cfgs, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t a_base = base, b_base = base + ROWS * 128, bar_ld = b_base + N * 128, bar_mma = bar_ld + 8, slot = bar_mma + 8;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(bar_ld, 1); mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (tid == 0) {
        mbar_expect_tx(bar_ld, ROWS * 128 + N * 128);
        tma_load_2d(a_base, &tmA, bar_ld, 0, 0);
        tma_load_2d(b_base, &tmB, bar_ld, 0, 0);
    }
    mbar_wait(bar_ld, 0);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t phase = 0;
    for (int ci = 0; ci < cfgs.n; ++ci) {
        const Cfg c = cfgs.c[ci];
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t start = a_base + (uint32_t)c.shift * 128u;
            const uint32_t bo = c.bo_mode ? ((start >> 7) & 7u) : 0u;
            for (int k = 0; k < 4; ++k)
                umma(tmem, make_desc(start, (uint32_t)c.sbo, bo) + (uint64_t)(2 * k), make_desc(b_base, 1024, 0) + (uint64_t)(2 * k), idesc, k != 0);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma) : "memory");
        }
        __syncwarp();
        mbar_wait(bar_mma, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr + (uint32_t)c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) out[((size_t)ci * 128 + tid) * N + c0 + j] = __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fp;
    std::vector<float> A(ROWS * 64), B(N * 64);
    std::vector<__nv_bfloat16> Ab(ROWS * 64), Bb(N * 64);
    srand(1);
    for (size_t i = 0; i < A.size(); ++i) { A[i] = (float)(rand() % 15 - 7); Ab[i] = __float2bfloat16(A[i]); }
    for (size_t i = 0; i < B.size(); ++i) { B[i] = (float)(rand() % 7 - 3); Bb[i] = __float2bfloat16(B[i]); }
    __nv_bfloat16 *dA, *dB; float* dOut;
    CK(cudaMalloc(&dA, Ab.size() * 2)); CK(cudaMalloc(&dB, Bb.size() * 2));
    CK(cudaMemcpy(dA, Ab.data(), Ab.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bb.data(), Bb.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[2] = {64, ROWS}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, ROWS}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode A failed %d\n", r); return 2; }
    }
    {
        cuuint64_t dims[2] = {64, N}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, N}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode B failed %d\n", r); return 2; }
    }
    Cfgs cfgs; cfgs.n = 0;
    const int shifts[] = {0, 1, 2, 3, 5, 8, 9, 19};
    const int sbos[] = {1024, 1152, 1280, 2304};
    for (int bo = 0; bo < 2; ++bo)
        for (int sb : sbos)
            for (int sh : shifts) {
                if (sh + 15 * (sb / 128) + 7 >= ROWS) continue;
                if (cfgs.n < NCFG_MAX) cfgs.c[cfgs.n++] = Cfg{sh, sb, bo};
            }
    CK(cudaMalloc(&dOut, (size_t)cfgs.n * 128 * N * 4));
    const size_t smem = ROWS * 128 + N * 128 + 1024 + 64;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<1, 128, smem>>>(tmA, tmB, cfgs, dOut);
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)cfgs.n * 128 * N);
    CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
    int all_abs_ok = 1;
    for (int ci = 0; ci < cfgs.n; ++ci) {
        const Cfg c = cfgs.c[ci];
        int bad_abs = 0, bad_rel = 0;
        for (int m = 0; m < 128; ++m) {
            const int rho = c.shift + (m / 8) * (c.sbo / 128) + (m % 8);
            for (int n = 0; n < N; ++n) {
                float e_abs = 0, e_rel = 0;
                for (int k = 0; k < 64; ++k) {
                    e_abs += A[rho * 64 + k] * B[n * 64 + k];
                    // hypothesis "relative": hardware XORs the chunk with (m & 7) instead of the address bits
                    const int chunk = k / 8, kk = k % 8;
                    const int src_chunk = chunk ^ (m & 7) ^ (rho & 7);
                    e_rel += A[rho * 64 + src_chunk * 8 + kk] * B[n * 64 + k];
                }
                const float g = out[((size_t)ci * 128 + m) * N + n];
                if (g != e_abs) ++bad_abs;
                if (g != e_rel) ++bad_rel;
            }
        }
        printf("shift %2d sbo %4d base_off_mode %d : mismatches absolute-address model %5d, relative model %5d  %s\n", c.shift, c.sbo, c.bo_mode, bad_abs, bad_rel,
               bad_abs == 0 ? "ABS-OK" : (bad_rel == 0 ? "REL-OK" : "NEITHER"));
        if (c.bo_mode == 0 && bad_abs) all_abs_ok = 0;
    }
    printf("RESULT: shifted/strided SWIZZLE_128B descriptors with base_offset=0 follow the absolute-address model: %s\n", all_abs_ok ? "YES" : "NO");
    return 0;
}
