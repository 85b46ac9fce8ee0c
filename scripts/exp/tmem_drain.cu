// Hardware experiment (sm_100a): how fast can the epilogue drain a tcgen05 accumulator, and what does each way of writing
// the result out cost -- alone and while the tensor pipe is busy?
//   reader warps (4 / 8 / 12; warp % 4 = TMEM lane quadrant) loop over "column groups" of 64 fp32 columns:
//     mode 0  tcgen05.ld only (2 x 32x32b.x32 + wait::ld)
//     mode 1  + convert to bf16 + 8 x st.shared.v4 into a 128-byte-swizzled staging row (the conv kernel's epilogue)
//     mode 2  + convert to bf16 + 8 x st.global.v4: each thread writes its pixel's 128 contiguous bytes (row pitch 1152 B)
//     mode 3  mode 1 + one TMA-store-like bulk copy per group (cp.async.bulk shared -> global of the 16 KB tile, by one
//             thread, waiting for the previous one's smem reads first: single staging buffer) ; mode 4: two staging buffers
//     mode 5  mode 1 per WARP (each warp stages its own 32 rows = 4 KB, __syncwarp only) + coalesced copy-out by the LSU:
//             8 x (ld.shared.v4 + st.global.v4), every instruction moving 4 rows x 128 B = four full 128-byte lines
//   mma = 1: warp 12 issues back-to-back tcgen05.mma (M=128, N=256, K=16, operands in shared memory) the whole time.
// Reports cycles per group per reader team (4 warps = 128 rows), the implied TMEM read rate in B/clk/SM, and the MMA rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../attention_based_tbn_b200/csrc -o bin/tmem_drain tmem_drain.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

using namespace tbn;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Cfg {
    int readers;  // warps
    int mode, mma, iters;
    __nv_bfloat16* out;  // per CTA: 128 rows x 576 bf16
};

__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(448, 1) kern(Cfg c, unsigned long long* cycles, unsigned long long* mma_cycles) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t a_base = base;                    // 16 KB
    const uint32_t b_base = base + 16384u;           // 32 KB
    const uint32_t stage0 = b_base + 32768u;         // 3 teams x 2 x 16 KB
    const uint32_t bars = stage0 + 6u * 16384u;
    const uint32_t slot = bars + 64u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t o = threadIdx.x * 16u; o < 16384u + 32768u; o += 448u * 16u)
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(0u) : "memory");
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bars + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async_smem();
    if (warp == 13) tmem_alloc(slot, 512u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
    volatile int* stop_ptr = reinterpret_cast<volatile int*>(raw + (bars + 56u - smem_u32(raw)));  // (dynamic smem: the kernel
    if (threadIdx.x == 0) *stop_ptr = 0;                                                            //  opts in to all 227 KB)
    __syncthreads();

    if (warp == 12) {
        if (c.mma) {
            const uint32_t idesc = umma_idesc_bf16(256);
            const uint64_t hi = umma_desc_sw128_hi(1024u);
            const uint64_t adesc = hi | (uint64_t)((a_base & 0x3FFFFu) >> 4);
            const uint64_t bdesc = hi | (uint64_t)((b_base & 0x3FFFFu) >> 4);
            const long long t0 = clock64();
            int n = 0;
            uint32_t ph = 0;
            while (!*stop_ptr) {
                if (elect_one()) {
                    for (int j = 0; j < 8; ++j) {
                        umma_bf16(tmem_base + 256u, adesc, bdesc, idesc, 1u);
                        umma_bf16(tmem_base + 256u, adesc + 2, bdesc + 2, idesc, 1u);
                        umma_bf16(tmem_base + 256u, adesc + 4, bdesc + 4, idesc, 1u);
                        umma_bf16(tmem_base + 256u, adesc + 6, bdesc + 6, idesc, 1u);
                    }
                    umma_commit(bars);
                }
                __syncwarp();
                mbar_wait(bars, ph);  // keep at most 32 MMAs in flight
                ph ^= 1u;
                n += 32;
            }
            if (lane == 0) mma_cycles[blockIdx.x] = (unsigned long long)((clock64() - t0) * 1000 / (n > 0 ? n : 1));
        }
    } else if (warp < c.readers) {
        const int team = warp >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t sw = (uint32_t)row & 7u;
        __nv_bfloat16* grow = c.out + ((size_t)blockIdx.x * 128 + row) * 576;
        uint32_t acc = 0;
        bool pending = false;
        const bool leader = q == 0 && lane == 0;
        const long long t0 = clock64();
        for (int i = 0; i < c.iters; ++i) {
            const uint32_t col = (uint32_t)((i * 64 + team * 64) & 192);  // accumulator columns 0..255 (the MMA uses 256..511)
            uint32_t v[64];
            tmem_ld32(taddr + col, v);
            tmem_ld32(taddr + col + 32, v + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c.mode == 0) {
#pragma unroll
                for (int j = 0; j < 64; ++j) acc ^= v[j];
                continue;
            }
            uint32_t o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[2 * j]), 0.f), fmaxf(__uint_as_float(v[2 * j + 1]), 0.f));
                o[j] = *reinterpret_cast<uint32_t*>(&h2);
            }
            if (c.mode == 2) {
                uint4* g = reinterpret_cast<uint4*>(grow + (i & 7) * 64);
#pragma unroll
                for (int j = 0; j < 8; ++j) g[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                continue;
            }
            if (c.mode == 5) {
                const uint32_t wstage = stage0 + (uint32_t)warp * 4096u + (uint32_t)(i & 1) * 12u * 4096u;  // 2 buffers per warp
                const uint32_t srow5 = wstage + (uint32_t)lane * 128u;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(srow5 + ((((uint32_t)j) ^ ((uint32_t)lane & 7u)) << 4)),
                                 "r"(o[4 * j]), "r"(o[4 * j + 1]), "r"(o[4 * j + 2]), "r"(o[4 * j + 3])
                                 : "memory");
                __syncwarp();
                const uint32_t ch = (uint32_t)lane & 7u, r0 = (uint32_t)lane >> 3;
                __nv_bfloat16* gbase = c.out + ((size_t)blockIdx.x * 128 + q * 32) * 576 + (i & 7) * 64 + ch * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t r = r0 + 4u * (uint32_t)j;
                    uint4 t;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
                                 : "r"(wstage + r * 128u + ((ch ^ (r & 7u)) << 4)));
                    *reinterpret_cast<uint4*>(gbase + (size_t)r * 576) = t;
                }
                __syncwarp();
                continue;
            }
            const uint32_t nbuf = c.mode == 4 ? 2u : 1u;
            const uint32_t stage = stage0 + ((uint32_t)team * 2u + ((uint32_t)i & (nbuf - 1u))) * 16384u;
            if (c.mode >= 3) {
                if (leader && pending) {
                    if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory");
            }
            const uint32_t srow = stage + (uint32_t)row * 128u;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((((uint32_t)j) ^ sw) << 4)), "r"(o[4 * j]),
                             "r"(o[4 * j + 1]), "r"(o[4 * j + 2]), "r"(o[4 * j + 3])
                             : "memory");
            if (c.mode >= 3) {
                fence_proxy_async_smem();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory");
                if (leader) {
                    bulk_store(c.out + (size_t)blockIdx.x * 128 * 576 + (size_t)(i & 7) * 8192, stage, 16384u);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    pending = true;
                }
            }
        }
        if (leader && pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        const long long t1 = clock64();
        if (acc == 0x12345u) c.out[0] = __float2bfloat16(1.f);
        if (lane == 0 && q == 0) cycles[blockIdx.x * 3 + team] = (unsigned long long)(t1 - t0);
        asm volatile("bar.sync 8, %0;" ::"r"(c.readers * 32) : "memory");
        if (warp == 0 && lane == 0) *stop_ptr = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 13) tmem_dealloc(tmem_base, 512u);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long *dc, *dm;
    CK(cudaMalloc(&dc, 148 * 3 * 8));
    CK(cudaMalloc(&dm, 148 * 8));
    __nv_bfloat16* out;
    CK(cudaMalloc(&out, (size_t)148 * 128 * 576 * 2));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const size_t smem = 1024 + 16384 + 32768 + 6 * 16384 + 256;
    printf("%-8s %-34s %-4s | cycles / 64-col group / team | TMEM read B/clk/SM | out B/clk/SM | cycles per N=256 MMA\n", "readers", "mode", "mma");
    const char* names[] = {"ld only", "ld + cvt + st.shared (swizzled)", "ld + cvt + st.global.v4 (direct)", "ld + cvt + sts + bulk store, 1 buf",
                           "ld + cvt + sts + bulk store, 2 bufs", "ld + cvt + sts + lds + coalesced stg"};
    for (int mma = 0; mma < 2; ++mma)
        for (int mode = 0; mode < 6; ++mode)
            for (int readers : {4, 8, 12}) {
                Cfg c;
                c.readers = readers; c.mode = mode; c.mma = mma; c.iters = 2000; c.out = out;
                CK(cudaMemset(dm, 0, 148 * 8));
                kern<<<sms, 448, smem>>>(c, dc, dm);
                CK(cudaDeviceSynchronize());
                kern<<<sms, 448, smem>>>(c, dc, dm);
                CK(cudaDeviceSynchronize());
                unsigned long long hc[148 * 3], hm[148];
                CK(cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(hm, dm, sizeof(hm), cudaMemcpyDeviceToHost));
                double avg = 0, am = 0;
                const int teams = readers / 4;
                for (int i = 0; i < sms; ++i) {
                    double worst = 0;
                    for (int t = 0; t < teams; ++t) worst = hc[i * 3 + t] > worst ? (double)hc[i * 3 + t] : worst;
                    avg += worst;
                    am += (double)hm[i] / 1000.0;
                }
                avg /= sms; am /= sms;
                const double per_group = avg / c.iters;
                printf("%-8d %-34s %-4d | %28.1f | %18.1f | %12.1f | %8.1f\n", readers, names[mode], mma, per_group,
                       teams * 128.0 * 64 * 4 / per_group, mode ? teams * 128.0 * 64 * 2 / per_group : 0.0, am);
            }
    return 0;
}
