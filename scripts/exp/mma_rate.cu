// Hardware experiment (sm_100a): what does one tcgen05.mma (kind::f16, bf16, K = 16) cost as a function of N, M, the
// number of issuing warps per CTA, the CTAs per SM and the source of A (shared memory descriptor vs tensor memory)?
// No data movement at all: operands are whatever shared / tensor memory holds; every issuing warp runs the loop of the
// conv kernel's MMA role (4 MMAs over a 64-channel k-block, one tcgen05.commit per k-block) and the CTA reports
// clock64 cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../attention_based_tbn_b200/csrc -o bin/mma_rate mma_rate.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

using namespace tbn;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// A operand from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// shared memory -> tensor memory copy of 128 lanes x 256 bits (one K = 16 slice of bf16 A) by the tensor-core proxy
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

struct Cfg {
    int m, n, issuers, a_tmem, iters, tmem_cols, same_acc;  // a_tmem: 0 smem, 1 tmem, 2 tmem filled by tcgen05.cp per MMA
};

__global__ void __launch_bounds__(128) kern(Cfg c, unsigned long long* cycles) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t a_base = base;                   // 2 stages x 16 KB
    const uint32_t b_base = base + 2u * 16384u;     // 2 stages x 32 KB
    const uint32_t bars = b_base + 2u * 32768u;     // per issuer: k-block barrier, final barrier
    const uint32_t slot = bars + 64u;
    const int warp = threadIdx.x >> 5;
    for (uint32_t o = threadIdx.x * 16u; o < 2u * 16384u + 2u * 32768u; o += 128u * 16u)
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(0u) : "memory");
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(bars + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(slot, (uint32_t)c.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
    if (warp < c.issuers) {
        const uint32_t idesc = idesc_bf16(c.m, c.n);
        const uint32_t ncol = (uint32_t)c.n;
        const uint32_t tmem_d = tmem_base + (c.same_acc ? 0u : (uint32_t)warp * ncol);
        const uint32_t tmem_a = tmem_base + (uint32_t)c.tmem_cols - 32u;  // 128 lanes x 32 columns = one 64-channel k-block
        const uint64_t hi = umma_desc_sw128_hi(1024u);
        const uint32_t kbar = bars + 16u * warp, fbar = kbar + 8u;
        const long long t0 = clock64();
        uint32_t accumulate = 0;
        for (int i = 0; i < c.iters; ++i) {
            const uint32_t st = (uint32_t)i & 1u;
            const uint64_t adesc = hi | (uint64_t)(((a_base + st * 16384u) & 0x3FFFFu) >> 4);
            const uint64_t bdesc = hi | (uint64_t)(((b_base + st * 32768u) & 0x3FFFFu) >> 4);
            if (elect_one()) {
                if (c.a_tmem == 2) {
                    // operand staged smem -> TMEM right before its MMA, two 32-column buffers alternating per k-block
                    const uint32_t ta = tmem_a - st * 32u;
                    tmem_cp_128x256b(ta, adesc);
                    tmem_cp_128x256b(ta + 8, adesc + 2);
                    tmem_cp_128x256b(ta + 16, adesc + 4);
                    tmem_cp_128x256b(ta + 24, adesc + 6);
                    umma_bf16_ts(tmem_d, ta, bdesc, idesc, accumulate);
                    umma_bf16_ts(tmem_d, ta + 8, bdesc + 2, idesc, 1u);
                    umma_bf16_ts(tmem_d, ta + 16, bdesc + 4, idesc, 1u);
                    umma_bf16_ts(tmem_d, ta + 24, bdesc + 6, idesc, 1u);
                } else if (c.a_tmem) {
                    umma_bf16_ts(tmem_d, tmem_a, bdesc, idesc, accumulate);
                    umma_bf16_ts(tmem_d, tmem_a + 8, bdesc + 2, idesc, 1u);
                    umma_bf16_ts(tmem_d, tmem_a + 16, bdesc + 4, idesc, 1u);
                    umma_bf16_ts(tmem_d, tmem_a + 24, bdesc + 6, idesc, 1u);
                } else {
                    umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
                    umma_bf16(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
                    umma_bf16(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
                    umma_bf16(tmem_d, adesc + 6, bdesc + 6, idesc, 1u);
                }
                umma_commit(kbar);
            }
            accumulate = 1u;
            __syncwarp();
        }
        if (elect_one()) umma_commit(fbar);
        __syncwarp();
        mbar_wait(fbar, 0);
        tc_fence_after();
        if (warp == 0 && (threadIdx.x & 31) == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)c.tmem_cols);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long* dc;
    CK(cudaMalloc(&dc, 2 * 148 * 8));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const size_t smem_need = 2 * 16384 + 2 * 32768 + 1024 + 256;
    printf("%-4s %-4s %-8s %-6s %-6s %-9s | cycles per MMA per CTA | per SM | pipe floor N/2 (M=128)\n", "M", "N", "issuers", "A", "CTA/SM",
           "same acc");
    const int ns[] = {32, 64, 96, 128, 160, 192, 256};
    for (int a_tmem = 0; a_tmem < 3; ++a_tmem)
        for (int m : {128, 64})
            for (int ctas : {1, 2})
                for (int issuers : {1, 2})
                    for (int same_acc : {0, 1})
                        for (int n : ns) {
                            if (same_acc && issuers == 1) continue;
                            Cfg c;
                            c.m = m; c.n = n; c.issuers = issuers; c.a_tmem = a_tmem; c.iters = 2000; c.same_acc = same_acc;
                            c.tmem_cols = ctas == 1 ? 512 : 256;
                            if (issuers * n + 64 > c.tmem_cols) continue;
                            if (a_tmem == 2 && (m != 128 || same_acc || issuers > 1)) continue;
                            // one CTA per SM: ask for more than half of the shared memory
                            const size_t smem = ctas == 1 ? 120 * 1024 : smem_need;
                            kern<<<sms * ctas, 128, smem>>>(c, dc);
                            CK(cudaDeviceSynchronize());
                            kern<<<sms * ctas, 128, smem>>>(c, dc);
                            CK(cudaDeviceSynchronize());
                            unsigned long long hc[2 * 148];
                            CK(cudaMemcpy(hc, dc, sizeof(unsigned long long) * sms * ctas, cudaMemcpyDeviceToHost));
                            double avg = 0;
                            for (int i = 0; i < sms * ctas; ++i) avg += (double)hc[i];
                            avg /= sms * ctas;
                            const double per_cta = avg / (c.iters * 4.0);
                            printf("%-4d %-4d %-8d %-6s %-6d %-9d | %8.1f | %8.1f | %5d\n", m, n, issuers, a_tmem == 2 ? "cp" : a_tmem ? "tmem" : "smem", ctas, same_acc,
                                   per_cta, per_cta / (issuers * ctas), n / 2);
                        }
    return 0;
}
