// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation).
//
//   D[M = pixels, N = Cout] = A[M, K = taps x Cin] * W[K, N]   (+ bias, ReLU, channel-offset store)
//
// Persistent, warp-specialised kernel: one CTA per SM loops over output tiles (128 pixels x BN channels).
//   warp 0   TMA producer (one elected lane).  A is never materialised in HBM:
//              * "tap" mode : for every (64-channel slice, filter tap) ONE 4-D TMA box {64 ch, BW, BH, BNI} of the
//                NHWC activation at the shifted pixel coordinate; out-of-bounds pixels (= conv padding) are
//                zero-filled by the TMA unit.  1x1 convs flatten (n,h,w) so that every tile is full.
//              * "halo" mode (3x3): ONE box {64 ch, 8+2, 16+2, 1} per 64-channel slice; the nine taps are nine
//                row-shifted views of that halo tile.  tcgen05 applies the 128-byte swizzle to absolute shared
//                memory address bits, so a K-major SWIZZLE_128B descriptor may start at any 128-byte row and use
//                any stride between its 8-row groups (measured: scripts/exp/umma_shift.cu) -- a tap is just
//                start += (r*10 + s)*128, stride-byte-offset = 10*128.  A is read from L2 once instead of 9 times.
//            W is pre-packed [Cout][tap][Cin padded to 64] bf16 (BatchNorm scale folded in), one 2-D box {64, BN} per
//            (slice, tap).  Boxes land in shared memory in the K-major SWIZZLE_128B operand layout of tcgen05.mma.
//   warp 1   MMA issuer: tcgen05.mma (M=128, N=BN<=256, K=16) into one of two TMEM accumulators;
//            tcgen05.commit releases the smem stages / publishes the accumulator through mbarriers.
//   warps 2+ epilogue, one to three teams of 4 warps (warp%4 = TMEM lane quadrant): tcgen05.ld 64 columns,
//            + folded bias, ReLU, convert, write 128 rows x 128 B into a SWIZZLE_128B staging tile (conflict-free),
//            then ONE thread issues a TMA tensor store of that tile into the channel slice of the block's concat
//            buffer (per-destination tensor maps clip partial tiles and segment ends, so there is no per-row
//            address arithmetic at all).  The epilogue of tile i overlaps the main loop of tile i+1
//            (double-buffered accumulator).
// Variants chosen per launch by the host code below: two CTAs per SM for N tiles <= 128 columns; CTA pairs
// (tcgen05 cta_group::2, template CTA2: 256 pixels x BN per pair, each CTA stages half of the weight tile) for wide
// tap-mode 3x3 convs; and, for conv2_3x3 (one K slice), CTA pairs that keep their half of all nine taps resident in
// shared memory for the life of the CTA, so that only the activation halo box is loaded per tile.
// All per-k-block loops are division-free: at ~100 instructions per k-block a single producer thread is slower
// than the tensor pipe (measured), so stage/phase counters are incremental and tile decoding is per tile.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

#define TBN_MAX_SEGS 4
#define TBN_MAX_GROUPS 264  // 8 attention heads x 1024 fp32 columns / 32 columns per group = 256
#define TBN_MAX_NTILES 126

namespace tbn {

constexpr int EPI_STAGE_BYTES = 128 * 128;        // staging tile per team: 128 rows x 128 B, SWIZZLE_128B
constexpr int HALO_BW = 8, HALO_BH = 16;          // halo-mode output tile: 8 x 16 pixels of one image
constexpr int HALO_PITCH = HALO_BW + 2;           // halo tile row pitch in pixels
constexpr int HALO_ROWS = HALO_PITCH * (HALO_BH + 2);  // 180 pixel rows of 128 B

struct TcParams {
    int n, h, w;             // image batch / spatial size (output == input, stride 1)
    int bw, bh, bni;         // output pixel tile (bw*bh*bni <= 128)
    int tiles_w, tiles_h;    // pixel tiles per image row / column (tiles over n implied)
    int n_tiles;             // tiles over Cout
    int total_tiles;         // m-tiles * n_tiles
    unsigned mg_nt, mg_tw, mg_th;  // floor(2^32 / d) for d = n_tiles, tiles_w, tiles_h (division-free tile decode)
    int cin;                 // input channels actually multiplied (multiple of 16)
    int cin_pad;             // cin rounded up to 64 (weight packing pitch per tap)
    int cchunks;             // cin_pad / 64
    int nk16_last;           // K=16 steps in the last 64-channel slice (zero-padded tail is skipped)
    int ksz, pad;            // filter size (1 or 3), padding
    int stride;              // 1, or 2 (3x3 tap mode: TMA boxes with element stride 2)
    int cout, bn_tile;
    int out_f32, bias_period;
    int halo;                // 1: halo mode
    int b_resident;          // 1: the CTA's whole weight tile (all taps) is loaded once and stays in shared memory
    int cta2_mul;            // 2 in CTA-pair mode (cta_group::2), else 1
    int ksplit, cc_per_split;  // split-K (GEMM mode only): tile index carries the K range, partial sums go to image `split`
    unsigned mg_ks;
    int a_stages, b_stages;  // tap mode: a_stages == b_stages (lock-step)
    unsigned a_stage_bytes, b_stage_bytes, a_tx_bytes;  // (b_stage_bytes: slot pitch, shared by the problems of a group)
    unsigned b_tx_bytes;     // bytes of this problem's weight box
    unsigned tmem_cols;      // per accumulator
    int epi_teams;           // 1 .. 3 teams of 4 epilogue warps
    int out_n;               // images of the output tensor (= n, or the number of K splits in split-K mode)
    int dbg;                 // ablation bits (tbn_set_tuning("tc_debug")): 1 no stores, 2 no epilogue, 4 no MMA,
                             // 8 MMA warp ignores the "full" barriers, 16 producer issues no TMA (8|16: MMA issue loop alone)
    unsigned long long* trace;  // diagnostic: per-role clock64 stamps of CTAs 0..3 (null = off)
    const float* bias;
    // output segments: columns [seg_start[s], seg_start[s+1]) go to destination s (tensor map OutMaps::m[s])
    int nseg;
    int seg_start[TBN_MAX_SEGS];
    int seg_flags[TBN_MAX_SEGS];  // TBN_SEG_RELU | TBN_SEG_BIAS
    // column groups: the unit of the epilogue (<= 128 B of output per row, never straddling a destination, so that
    // every TMA store has non-negative coordinates -- negative store coordinates fault on sm_100a, measured)
    short grp_col[TBN_MAX_GROUPS + 1];      // first output column of group g; [ngroups] = cout
    signed char grp_seg[TBN_MAX_GROUPS];    // destination of group g
    unsigned short tile_grp[TBN_MAX_NTILES + 2];  // N tile t = groups [tile_grp[t], tile_grp[t+1])
};

struct OutMaps {
    CUtensorMap m[TBN_MAX_SEGS];  // one store map per destination; its channel extent = the segment's columns
};

// One convolution / GEMM problem of a launch.  A GROUPED launch (TcGroup<NP>, NP > 1) walks the concatenated tile lists
// of several problems -- the same layer of the RGB / flow / audio backbones, or the independent 3x3 convs of one
// inception block -- with one persistent grid: a launch then has enough tiles to fill whole waves of the 148 SMs (the
// 7x7 / 8x13 maps of inception 5a / 5b alone have 1.04-2.6 waves) and pays the pipeline fill / drain once.  All
// problems of a group share the shared-memory / tensor-memory plan (stage sizes and counts, epilogue teams, accumulator
// columns: the host takes the maximum over the problems) and the kernel variant (HALO, CTA2).
struct TcProblem {
    CUtensorMap tmA, tmB;
    OutMaps tmY;
    TcParams p;
};
template <int NP>
struct TcGroup {
    int nprob, total_tiles;
    int tile_end[NP];  // cumulative tile counts: problem i owns tiles [tile_end[i-1], tile_end[i])
    TcProblem pr[NP];
};
#define TBN_MAX_PROBS 6
template <int NP>
__device__ __forceinline__ int find_problem(const TcGroup<NP>& g, int tile, int& local) {
    int pi = 0;
    local = tile;
    if (NP > 1) {
#pragma unroll
        for (int i = 0; i < NP - 1; ++i)
            if (i + 1 < g.nprob && tile >= g.tile_end[i]) {
                pi = i + 1;
                local = tile - g.tile_end[i];
            }
    }
    return pi;
}

// trace layout per CTA (TRACE_CTAS CTAs): [0] start, [1] end, then 4 arrays of TRACE_N: producer issue, mma full-wait done,
// mma commit, epilogue (2 per tile: start, end)
constexpr int TRACE_N = 256;
constexpr int TRACE_CTAS = 4;
constexpr int TRACE_STRIDE = 8 + 4 * TRACE_N;
__device__ __forceinline__ void trace_put(const TcParams& p, int arr, uint32_t idx) {
    if (p.trace != nullptr && blockIdx.x < TRACE_CTAS && idx < (uint32_t)TRACE_N && (threadIdx.x & 31) == 0)
        p.trace[(size_t)blockIdx.x * TRACE_STRIDE + 8 + arr * TRACE_N + idx] = (unsigned long long)clock64();
}

// n = q*d + r with mg = floor(2^32 / d): the estimate umulhi(n, mg) is q or q-1, one correction step makes it exact
__device__ __forceinline__ void fast_divmod(uint32_t n, uint32_t d, uint32_t mg, uint32_t& q, uint32_t& r) {
    q = __umulhi(n, mg);
    r = n - q * d;
    if (r >= d) { ++q; r -= d; }
}

// tile -> pixel tile origin and column groups.  In CTA-pair mode `tile` indexes a PAIR of pixel tiles and the CTA of
// rank r takes pixel tile 2*pair + r (a trailing odd tile falls outside the tensor: TMA zero-fills / clips it).
__device__ __forceinline__ void decode_tile(const TcParams& p, int tile, uint32_t pair_rank, int& w0, int& h0, int& n0, int& g0,
                                            int& g1, int& cc0, int& cc1) {
    uint32_t mt, n_tile, tw, th, tn, t = (uint32_t)tile, ks = 0;
    if (p.ksplit > 1) fast_divmod(t, (uint32_t)p.ksplit, p.mg_ks, t, ks);
    cc0 = (int)ks * p.cc_per_split;
    cc1 = cc0 + p.cc_per_split < p.cchunks ? cc0 + p.cc_per_split : p.cchunks;
    fast_divmod(t, (uint32_t)p.n_tiles, p.mg_nt, mt, n_tile);
    mt = mt * (uint32_t)p.cta2_mul + pair_rank;
    fast_divmod(mt, (uint32_t)p.tiles_w, p.mg_tw, mt, tw);
    fast_divmod(mt, (uint32_t)p.tiles_h, p.mg_th, tn, th);
    w0 = (int)tw * p.bw; h0 = (int)th * p.bh; n0 = (int)tn * p.bni;
    if (p.ksplit > 1) n0 = (int)ks;  // split-K (flattened GEMM: tn == 0): the partial sum goes to image `ks` of the workspace
    g0 = p.tile_grp[n_tile]; g1 = p.tile_grp[n_tile + 1];  // column groups of this N tile
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// 16 accumulator columns -> bias, ReLU, convert, staging row (128 B per row, 16-byte chunk c stored at c ^ (row & 7))
__device__ __forceinline__ void epi_chunk16(const TcParams& p, const uint32_t* v, int col, int flags, const float* brow,
                                            uint32_t stage_row, uint32_t sw, int chunk16) {
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
    if (flags & TBN_SEG_BIAS) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(brow + col + j));
            f[j + 0] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
        }
    }
    if (flags & TBN_SEG_RELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (p.out_f32) {  // 16 fp32 = 64 B = staging chunks 4*chunk16 .. +3
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + ((((uint32_t)(chunk16 * 4 + (j >> 2))) ^ sw) << 4)),
                         "f"(f[j]), "f"(f[j + 1]), "f"(f[j + 2]), "f"(f[j + 3])
                         : "memory");
    } else {          // 16 bf16 = 32 B = staging chunks 2*chunk16, +1
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            o[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + ((((uint32_t)(chunk16 * 2)) ^ sw) << 4)), "r"(o[0]),
                     "r"(o[1]), "r"(o[2]), "r"(o[3])
                     : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + ((((uint32_t)(chunk16 * 2 + 1)) ^ sw) << 4)), "r"(o[4]),
                     "r"(o[5]), "r"(o[6]), "r"(o[7])
                     : "memory");
    }
}

template <bool HALO, bool CTA2, int NP>
__global__ void __launch_bounds__(448, 1) conv_tc_kernel(const __grid_constant__ TcGroup<NP> g) {
    const TcParams& p = g.pr[0].p;  // shared plan: stages, teams, accumulator columns, debug switches
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
    const uint32_t a_smem = base;
    const uint32_t b_smem = a_smem + (uint32_t)p.a_stages * p.a_stage_bytes;
    const uint32_t epi_smem = b_smem + (uint32_t)p.b_stages * p.b_stage_bytes;
    const uint32_t bar_base = epi_smem + (uint32_t)p.epi_teams * EPI_STAGE_BYTES;
    const uint32_t a_full = bar_base, a_empty = a_full + 8u * p.a_stages;
    const uint32_t b_full = a_empty + 8u * p.a_stages, b_empty = b_full + 8u * p.b_stages;
    const uint32_t accf0 = b_empty + 8u * p.b_stages;  // accumulator full [2]
    const uint32_t acce0 = accf0 + 16u;                 // accumulator empty [2]
    const uint32_t tmem_slot = acce0 + 16u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Programmatic dependent launch: let the next kernel of the stream start its prologue (barrier init, TMEM allocation,
    // tensor-map prefetch) on SMs this grid has already left; this grid itself touches global memory only after
    // griddepcontrol.wait below, i.e. after the grids it depends on have completed and flushed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // CTA-pair mode: the two CTAs of a cluster compute 256 pixels x BN together (tcgen05 cta_group::2): each loads its own
    // 128-pixel A tile and HALF of the weight tile, the leader (rank 0) issues the MMAs for both, every CTA drains its own
    // TMEM.  Per CTA this halves the weight bytes through L2, TMA and -- the binding one -- the shared-memory port.
    const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
    const int unit = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // work-loop index of this CTA / pair
    const int nunits = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.a_stages; ++s) {
            mbar_init(a_full + 8u * s, 1);
            mbar_init(a_empty + 8u * s, 1);
        }
        for (int s = 0; s < p.b_stages; ++s) {
            mbar_init(b_full + 8u * s, 1);
            mbar_init(b_empty + 8u * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(accf0 + 8u * a, 1);
            mbar_init(acce0 + 8u * a, (CTA2 ? 8u : 4u) * p.epi_teams);  // one arrival per epilogue warp (of both CTAs)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < g.nprob; ++i) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&g.pr[i].tmA)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&g.pr[i].tmB)) : "memory");
            for (int s = 0; s < g.pr[i].p.nseg; ++s)
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&g.pr[i].tmY.m[s])) : "memory");
        }
    }
    if (warp == 1) {
        if (CTA2) tmem_alloc_2sm(tmem_slot, 2u * p.tmem_cols);
        else tmem_alloc(tmem_slot, 2u * p.tmem_cols);
    }
    tc_fence_before();
    __syncthreads();
    if (CTA2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    asm volatile("griddepcontrol.wait;" ::: "memory");  // everything above overlapped the previous kernel's tail
    if (p.trace != nullptr && blockIdx.x < TRACE_CTAS && threadIdx.x == 0)
        p.trace[(size_t)blockIdx.x * TRACE_STRIDE] = (unsigned long long)clock64();

    if (warp == 0) {
        {
            // ===== TMA producer (whole warp converged; the elected lane issues) =====
            uint32_t as = 0, aph = 0, bs = 0, bph = 0, kbt = 0;
            const uint32_t mul = CTA2 ? 2u : 1u;  // the leader's barrier collects the bytes of both CTAs
            const uint32_t a_full_ld = CTA2 ? mapa_shared(a_full, 0) : a_full;  // "full" barriers live in the leader
            const uint32_t b_full_ld = CTA2 ? mapa_shared(b_full, 0) : b_full;
            for (int tile = unit; tile < g.total_tiles; tile += nunits) {
                int ltile;
                const TcProblem& pr = g.pr[find_problem(g, tile, ltile)];
                const TcParams& p = pr.p;
                const CUtensorMap& tmA = pr.tmA;
                const CUtensorMap& tmB = pr.tmB;
                const uint32_t ab_tx = (p.a_tx_bytes + p.b_tx_bytes) * mul;
                int w0, h0, n0, g0, g1, cc0, cc1;
                decode_tile(p, ltile, rank, w0, h0, n0, g0, g1, cc0, cc1);
                if (p.ksplit > 1) n0 = 0;  // (loads: the A operand has a single image)
                int col0 = p.grp_col[g0];
                if (CTA2) col0 += (int)rank * ((p.grp_col[g1] - col0) >> 1);  // this CTA's half of the weight tile
                for (int cc = cc0; cc < cc1; ++cc) {
                    const int c0 = cc * 64;
                    if (HALO) {
                        mbar_wait(a_empty + 8u * as, aph ^ 1u);
                        if (!(p.dbg & 16) && elect_one()) {
                            if (rank == 0) mbar_expect_tx(a_full + 8u * as, p.a_tx_bytes * mul);
                            if (CTA2)
                                tma_load_4d_2sm(a_smem + as * p.a_stage_bytes, &tmA, a_full_ld + 8u * as, c0, w0 - 1, h0 - 1, n0);
                            else
                                tma_load_4d(a_smem + as * p.a_stage_bytes, &tmA, a_full + 8u * as, c0, w0 - 1, h0 - 1, n0);
                        }
                        if (++as == (uint32_t)p.a_stages) { as = 0; aph ^= 1u; }
                    }
                    if (p.b_resident && tile != unit) continue;  // resident weights: loaded with the CTA's first tile only
                    int kc = c0;  // weight K coordinate = tap * cin_pad + cc * 64
                    for (int r = 0; r < p.ksz; ++r)
                        for (int s = 0; s < p.ksz; ++s, kc += p.cin_pad) {
                            if (!p.b_resident) mbar_wait(b_empty + 8u * bs, bph ^ 1u);
                            trace_put(p, 0, kbt++);
                            if (!(p.dbg & 16) && elect_one()) {
                                if (rank == 0) mbar_expect_tx(b_full + 8u * bs, HALO ? p.b_tx_bytes * mul : ab_tx);
                                if (!HALO) {
                                    if (CTA2)
                                        tma_load_4d_2sm(a_smem + bs * p.a_stage_bytes, &tmA, b_full_ld + 8u * bs, c0,
                                                        w0 * p.stride + s - p.pad, h0 * p.stride + r - p.pad, n0);
                                    else
                                        tma_load_4d(a_smem + bs * p.a_stage_bytes, &tmA, b_full + 8u * bs, c0,
                                                    w0 * p.stride + s - p.pad, h0 * p.stride + r - p.pad, n0);
                                }
                                if (CTA2) tma_load_2d_2sm(b_smem + bs * p.b_stage_bytes, &tmB, b_full_ld + 8u * bs, kc, col0);
                                else tma_load_2d(b_smem + bs * p.b_stage_bytes, &tmB, b_full + 8u * bs, kc, col0);
                            }
                            if (++bs == (uint32_t)p.b_stages) { bs = 0; bph ^= 1u; }
                        }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (rank == 0) {
            // ===== MMA issuer (whole warp converged, elected lane issues; CTA-pair mode: the leader issues for both CTAs) =====
            const uint64_t adesc_hi = umma_desc_sw128_hi(HALO ? (uint32_t)HALO_PITCH * 128u : 1024u);
            const uint64_t bdesc_hi = umma_desc_sw128_hi(1024u);
            uint32_t as = 0, aph = 0, bs = 0, bph = 0, it = 0, kbt = 0;
            for (int tile = unit; tile < g.total_tiles; tile += nunits, ++it) {
                const uint32_t acc = it & 1u;
                int ltile;
                const TcParams& p = g.pr[find_problem(g, tile, ltile)].p;
                int w0, h0, n0, g0, g1, cc0, cc1;
                decode_tile(p, ltile, 0u, w0, h0, n0, g0, g1, cc0, cc1);
                const int tile_n = p.grp_col[g1] - p.grp_col[g0];  // N = this tile's columns
                const uint32_t idesc = CTA2 ? umma_idesc_bf16_2sm(tile_n) : umma_idesc_bf16(tile_n);
                mbar_wait(acce0 + 8u * acc, ((it >> 1) & 1u) ^ 1u);  // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * p.tmem_cols;
                uint32_t accumulate = 0;
                for (int cc = cc0; cc < cc1; ++cc) {
                    const int nk16 = (cc == p.cchunks - 1) ? p.nk16_last : 4;
                    uint32_t a_addr = 0;
                    if (HALO) {
                        if (!(p.dbg & 8)) mbar_wait(a_full + 8u * as, aph);
                        a_addr = a_smem + as * p.a_stage_bytes;
                    }
                    for (int r = 0; r < p.ksz; ++r)
                        for (int s = 0; s < p.ksz; ++s) {
                            if ((!p.b_resident || tile == unit) && !(p.dbg & 8)) mbar_wait(b_full + 8u * bs, p.b_resident ? 0u : bph);
                            tc_fence_after();
                            trace_put(p, 1, kbt);
                            const uint32_t a_tap = HALO ? a_addr + (uint32_t)(r * HALO_PITCH + s) * 128u
                                                        : a_smem + bs * p.a_stage_bytes;
                            const uint64_t adesc = adesc_hi | (uint64_t)((a_tap & 0x3FFFFu) >> 4);
                            const uint64_t bdesc = bdesc_hi | (uint64_t)(((b_smem + bs * p.b_stage_bytes) & 0x3FFFFu) >> 4);
                            if (elect_one()) {
                                if (!(p.dbg & 4)) {
                                    // full 64-channel slice: four independent issues (no loop-carried descriptor chain)
                                    if (nk16 == 4) {
                                        if (CTA2) {
                                            umma_bf16_2sm(tmem_d, adesc, bdesc, idesc, accumulate);
                                            umma_bf16_2sm(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
                                            umma_bf16_2sm(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
                                            umma_bf16_2sm(tmem_d, adesc + 6, bdesc + 6, idesc, 1u);
                                        } else if (p.dbg & 32) {  // timing experiment: two independent accumulator chains
                                            const uint32_t d1 = tmem_base + (acc ^ 1u) * p.tmem_cols;
                                            umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
                                            umma_bf16(d1, adesc + 2, bdesc + 2, idesc, accumulate);
                                            umma_bf16(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
                                            umma_bf16(d1, adesc + 6, bdesc + 6, idesc, 1u);
                                        } else {
                                            umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
                                            umma_bf16(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
                                            umma_bf16(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
                                            umma_bf16(tmem_d, adesc + 6, bdesc + 6, idesc, 1u);
                                        }
                                    } else {
                                        uint32_t accum = accumulate;
                                        for (int k16 = 0; k16 < nk16; ++k16) {
                                            if (CTA2)
                                                umma_bf16_2sm(tmem_d, adesc + (uint64_t)(2 * k16), bdesc + (uint64_t)(2 * k16), idesc,
                                                              accum);
                                            else
                                                umma_bf16(tmem_d, adesc + (uint64_t)(2 * k16), bdesc + (uint64_t)(2 * k16), idesc,
                                                          accum);
                                            accum = 1u;
                                        }
                                    }
                                }
                                // frees the smem stage (in both CTAs) once these MMAs retire
                                if (p.b_resident) {
                                } else if (CTA2) {
                                    umma_commit_2sm(b_empty + 8u * bs);
                                } else {
                                    umma_commit(b_empty + 8u * bs);
                                }
                            }
                            accumulate = 1u;
                            trace_put(p, 2, kbt++);
                            if (++bs == (uint32_t)p.b_stages) { bs = 0; bph ^= 1u; }
                        }
                    if (HALO) {
                        if (elect_one()) {
                            if (CTA2) umma_commit_2sm(a_empty + 8u * as);
                            else umma_commit(a_empty + 8u * as);
                        }
                        if (++as == (uint32_t)p.a_stages) { as = 0; aph ^= 1u; }
                    }
                }
                // accumulator complete (both CTAs' epilogues wait on their own copy of this barrier)
                if (elect_one()) {
                    if (CTA2) umma_commit_2sm(accf0 + 8u * acc);
                    else umma_commit(accf0 + 8u * acc);
                }
            }
        }
        __syncwarp();
    } else if (warp < 2 + 4 * p.epi_teams) {
        // ===== epilogue: team = 4 warps covering the four TMEM lane quadrants (quadrant = warp % 4) =====
        const int team = (warp - 2) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;                  // accumulator row (= TMEM lane) = row of the staging tile
        const bool leader = ((warp - 2) & 3) == 0 && lane == 0;  // issues the team's TMA stores
        const uint32_t stage = epi_smem + (uint32_t)team * EPI_STAGE_BYTES;
        const uint32_t stage_row = stage + (uint32_t)row * 128u;
        const uint32_t sw = (uint32_t)row & 7u;
        uint32_t it = 0;
        bool store_pending = false;  // leader: a TMA store may still be reading the staging tile
        const uint32_t acce_ld = CTA2 ? mapa_shared(acce0, 0) : acce0;  // "accumulator empty" lives in the leader
        for (int tile = unit; tile < g.total_tiles; tile += nunits, ++it) {
            int ltile;
            const TcProblem& pr = g.pr[find_problem(g, tile, ltile)];
            const TcParams& p = pr.p;
            const OutMaps& tmY = pr.tmY;
            int w0, h0, n0, g0, g1, cc0, cc1;
            decode_tile(p, ltile, rank, w0, h0, n0, g0, g1, cc0, cc1);
            const int col0 = p.grp_col[g0];
            const uint32_t acc = it & 1u;
            const float* brow = p.bias;
            if (p.bias_period > 1) {  // per-position bias table (row index modulo the period)
                const int wi = row % p.bw, hi = (row / p.bw) % p.bh, ni = row / (p.bw * p.bh);
                const int mflat = ((n0 + ni) * p.h + (h0 + hi)) * p.w + (w0 + wi);
                brow += (long long)(mflat % p.bias_period) * p.cout;
            }
            mbar_wait(accf0 + 8u * acc, (it >> 1) & 1u);
            tc_fence_after();
            if (warp == 2 && lane == 0) trace_put(p, 3, 2 * it);
            const uint32_t taddr = tmem_base + acc * p.tmem_cols + ((uint32_t)(q * 32) << 16);
            const int gfirst = g0 + team;
            if (!(p.dbg & 2)) {
                for (int g = gfirst; g < g1; g += p.epi_teams) {
                    const int col = p.grp_col[g];
                    const int ncol = p.grp_col[g + 1] - col;   // 16 .. 64 (bf16) / 32 (fp32), multiple of 16
                    const int sg = p.grp_seg[g];
                    const int flags = p.seg_flags[sg];
                    // ---- TMEM -> registers (both halves in flight before the wait)
                    uint32_t v[64];
                    const bool second = ncol > 32;
                    if (ncol > 16) tmem_ld32(taddr + (uint32_t)(col - col0), v);  // never read past the group: the last
                    else tmem_ld16(taddr + (uint32_t)(col - col0), v);            // one may end at the TMEM allocation
                    if (ncol > 48) tmem_ld32(taddr + (uint32_t)(col - col0 + 32), v + 32);
                    else if (second) tmem_ld16(taddr + (uint32_t)(col - col0 + 32), v + 32);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (g + p.epi_teams >= g1) {  // last TMEM read of this team: hand the accumulator back early
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CTA2) mbar_arrive_cluster(acce_ld + 8u * acc);
                            else mbar_arrive(acce0 + 8u * acc);
                        }
                    }
                    // ---- the previous TMA store must have finished reading the staging tile
                    if (leader && store_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    named_bar_sync(1 + team, 128);
                    // ---- bias / ReLU / convert -> swizzled staging rows
                    epi_chunk16(p, v, col, flags, brow, stage_row, sw, 0);
                    if (ncol > 16) epi_chunk16(p, v + 16, col + 16, flags, brow, stage_row, sw, 1);
                    if (second) {
                        epi_chunk16(p, v + 32, col + 32, flags, brow, stage_row, sw, 2);
                        if (ncol > 48) epi_chunk16(p, v + 48, col + 48, flags, brow, stage_row, sw, 3);
                    }
                    fence_proxy_async_smem();       // generic-proxy writes -> visible to the TMA (async proxy)
                    named_bar_sync(1 + team, 128);
                    // ---- one thread stores the tile; the destination's tensor map clips partial tiles / short groups
                    if (leader && !(p.dbg & 1)) {
                        tma_store_4d(&tmY.m[sg], stage, col - p.seg_start[sg], w0, h0, n0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        store_pending = true;
                    }
                }
            }
            if ((p.dbg & 2) || gfirst >= g1) {  // nothing read from TMEM by this team: still release the accumulator
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CTA2) mbar_arrive_cluster(acce_ld + 8u * acc);
                    else mbar_arrive(acce0 + 8u * acc);
                }
            }
            if (warp == 2 && lane == 0) trace_put(p, 3, 2 * it + 1);
        }
        if (leader && store_pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
    }

    tc_fence_before();
    __syncthreads();
    if (CTA2) cluster_sync_all();  // neither CTA leaves (freeing smem / TMEM) while its peer may still use them
    if (p.trace != nullptr && blockIdx.x < TRACE_CTAS && threadIdx.x == 0)
        p.trace[(size_t)blockIdx.x * TRACE_STRIDE + 1] = (unsigned long long)clock64();
    if (warp == 1) {
        if (CTA2) tmem_dealloc_2sm(tmem_base, 2u * p.tmem_cols);
        else tmem_dealloc(tmem_base, 2u * p.tmem_cols);
    }
}

// ---- host side ----------------------------------------------------------------------------------
static int g_tc_halo = -1;  // -1: read TBN_TC_HALO on first use (default 1 = halo mode where it pays)
static int g_tc_debug = 0;
static int g_tc_bres = 2;    // resident weights for single-slice (cin <= 64) halo-mode 3x3 convs.  1: every such conv with N tiles
                             // of <= 96 columns (measured slower: those tiles are bound by the shared-memory operand floor of
                             // narrow MMAs); 2 (default): only as CTA PAIRS for 128 < cout <= 192, i.e. conv2_3x3 -- each CTA
                             // of a pair keeps its half of all nine taps (108 KB) for its whole life, the MMA keeps N = 192,
                             // and the per-tile TMA traffic drops from 239 KB to the 23 KB halo box: -12 % / -14 % (RGB / audio)
static int g_tc_pdl = 0;     // 1: programmatic dependent launch (prologue overlaps the previous kernel's tail); measured: no
                             // gain inside the captured multi-stream graph, so it is off by default
static int g_tc_max_bn = 256;  // widest N tile (columns); per-layer autotuning lowers it where narrower tiles win
static int g_tc_occ2 = 1;    // 1: two CTAs per SM for narrow N tiles
static int g_tc_cta2 = 3;    // CTA pairs (cta_group::2): 0 never, 1 for N tiles wider than 128 columns, 2 always, 3 (default) only
                             // for tap-mode 3x3 convs with wide N tiles -- the one class where pairs measured faster
                             // (2-6 % per layer, +1.2 % per step); elsewhere the cross-CTA epilogue handshake costs more
                             // than the halved weight traffic saves (fused 1x1 GEMMs: up to 28 % slower)
static int g_tc_teams3 = 1;  // 1: allow a third epilogue team where it shortens the longest team's work
static unsigned long long* g_tc_trace = nullptr;  // device buffer of TRACE_CTAS * TRACE_STRIDE stamps (diagnostic)

EncodeTiledFn get_tensormap_encoder() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Pixel box {bw, bh, bni} with bw*bh*bni <= 128 that minimises the number of 128-row tiles.
static long long pick_box(int n, int h, int w, int& bw, int& bh, int& bni) {
    long long best = -1;
    bw = bh = bni = 1;
    for (int cw = 1; cw <= w && cw <= 128; ++cw)
        for (int ch = 1; ch <= h && cw * ch <= 128; ++ch) {
            int cn = 128 / (cw * ch);
            if (cn > n) cn = n;
            const long long tiles = (long long)cdiv(w, cw) * cdiv(h, ch) * cdiv(n, cn);
            if (best < 0 || tiles < best || (tiles == best && cw * ch > bw * bh)) {
                best = tiles; bw = cw; bh = ch; bni = cn;
            }
        }
    return best;
}

struct OutSeg {
    void* y;
    int y_cs, y_co, cols, flags;
};

// One problem, planned: kernel parameters + tensor maps + what the launch has to agree on across a group.
struct ConvPlan {
    TcProblem pr;
    bool cta2;
    int ctas_per_sm, nt, bn;
    long long m_tiles;
    long long cost;  // per-tile work (k-blocks x columns): longest tiles first in a grouped launch
};

static int plan_conv_tc(ConvPlan& pl, const void* x, const void* wp, const float* bias, const OutSeg* segs, int nseg, int n, int h_in,
                        int w_in, int cin, int x_cs, int x_co, int k, int stride, int out_dtype, int bias_period, int ksplit,
                        bool grouped, bool group_fills_gpu = false) {
    TcParams& p = pl.pr.p;
    CUtensorMap& tmA = pl.pr.tmA;
    CUtensorMap& tmB = pl.pr.tmB;
    OutMaps& tmY = pl.pr.tmY;
    TBN_CHECK_ARG(stride == 1 || (stride == 2 && k == 3), "conv_tc: stride must be 1, or 2 with k = 3");
    const int h = stride == 1 ? h_in : (h_in - 1) / 2 + 1;  // output size (pad = k/2)
    const int w = stride == 1 ? w_in : (w_in - 1) / 2 + 1;
    TBN_CHECK_ARG(x && wp && bias && segs, "conv_tc: null pointer");
    TBN_CHECK_ARG(n > 0 && h_in > 0 && w_in > 0, "conv_tc: bad geometry");
    TBN_CHECK_ARG(k == 1 || k == 3, "conv_tc: k must be 1 or 3 (got %d); strided convs go through tbn_im2col", k);
    TBN_CHECK_ARG(cin % 16 == 0 && x_cs % 8 == 0 && x_co % 8 == 0, "conv_tc: cin %% 16, x_cs %% 8, x_co %% 8 required");
    TBN_CHECK_ARG(nseg >= 1 && nseg <= TBN_MAX_SEGS, "conv_tc: 1..%d output segments", TBN_MAX_SEGS);
    TBN_CHECK_ARG(out_dtype == TBN_BF16 || out_dtype == TBN_F32, "conv_tc: bad out_dtype");
    TBN_CHECK_ARG(bias_period >= 1, "conv_tc: bias_period >= 1");
    TBN_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 15) == 0, "conv_tc: pointers must be 16-byte aligned");
    TBN_CHECK_ARG((long long)n * h * w < (1LL << 31), "conv_tc: too many pixels");
    EncodeTiledFn encode = get_tensormap_encoder();
    TBN_CHECK_ARG(encode != nullptr, "conv_tc: cuTensorMapEncodeTiled entry point not available");

    int cout = 0;
    for (int s = 0; s < nseg; ++s) {
        TBN_CHECK_ARG(segs[s].y && ((uintptr_t)segs[s].y & 15) == 0, "conv_tc: output pointers must be 16-byte aligned");
        TBN_CHECK_ARG(segs[s].cols > 0 && segs[s].cols % 16 == 0 && segs[s].y_cs % 8 == 0 && segs[s].y_co % 8 == 0,
                      "conv_tc: cout %% 16, y_cs %% 8, y_co %% 8 required");
        p.seg_flags[s] = segs[s].flags;
        p.seg_start[s] = cout;
        cout += segs[s].cols;
    }
    for (int s = nseg; s < TBN_MAX_SEGS; ++s) {
        p.seg_flags[s] = 0; p.seg_start[s] = 0x7fffffff;
    }
    p.nseg = nseg;
    if (g_tc_halo < 0) {
        const char* e = getenv("TBN_TC_HALO");
        g_tc_halo = e ? atoi(e) : 1;
        if (const char* pe = getenv("TBN_TC_PDL")) g_tc_pdl = atoi(pe) ? 1 : 0;
        if (const char* ce = getenv("TBN_TC_CTA2")) g_tc_cta2 = atoi(ce);
    }
    // A 1x1 conv has no halo: flatten (n,h,w) into one pixel axis so that every tile is full.
    if (k == 1) {
        p.n = 1; p.h = 1; p.w = n * h * w;
    } else {
        p.n = n; p.h = h; p.w = w;
    }
    long long m_tiles = pick_box(p.n, p.h, p.w, p.bw, p.bh, p.bni);
    p.halo = 0;
    if (k == 3 && stride == 1 && g_tc_halo) {
        // halo mode reads A once instead of nine times; accept up to ~15% more (partly empty) tiles for it
        const long long halo_tiles = (long long)cdiv(w, HALO_BW) * cdiv(h, HALO_BH) * n;
        if (g_tc_halo == 2 || halo_tiles * 100 <= m_tiles * 115) {  // (2 = forced)
            p.halo = 1;
            p.bw = HALO_BW; p.bh = HALO_BH; p.bni = 1;
            m_tiles = halo_tiles;
        }
    }
    p.tiles_w = cdiv(p.w, p.bw);
    p.tiles_h = cdiv(p.h, p.bh);
    p.cin = cin;
    p.cin_pad = (cin + 63) / 64 * 64;
    p.cchunks = p.cin_pad / 64;
    p.nk16_last = (cin - (p.cchunks - 1) * 64) / 16;
    p.ksz = k;
    p.pad = k / 2;
    p.stride = stride;
    p.cout = cout;
    // Column groups (<= 128 B of output per pixel, never straddling a destination) and N tiles = runs of groups.
    // N tile: as wide as possible (<= 256 columns: A is then read once), narrowed while the grid underfills the GPU.
    const int gcols = out_dtype == TBN_F32 ? 32 : 64;
    const int sms = sm_count();
    int ngroups = 0;
    for (int sg = 0; sg < nseg; ++sg)
        for (int c = 0; c < segs[sg].cols; c += gcols) {
            TBN_CHECK_ARG(ngroups < TBN_MAX_GROUPS, "conv_tc: too many output column groups (cout %d)", cout);
            p.grp_col[ngroups] = (short)(p.seg_start[sg] + c);
            p.grp_seg[ngroups] = (signed char)sg;
            ++ngroups;
        }
    p.grp_col[ngroups] = (short)cout;
    auto partition = [&](int max_cols, int* max_tile_cols) {  // greedy: N tiles of at most max_cols columns
        int nt_ = 0, g = 0, widest = 0;
        while (g < ngroups) {
            if (nt_ < TBN_MAX_NTILES) p.tile_grp[nt_] = (unsigned short)g;
            const int start = p.grp_col[g];
            ++g;
            while (g < ngroups && p.grp_col[g + 1] - start <= max_cols) ++g;
            if (p.grp_col[g] - start > widest) widest = p.grp_col[g] - start;
            ++nt_;
        }
        if (nt_ <= TBN_MAX_NTILES) p.tile_grp[nt_] = (unsigned short)ngroups;
        if (max_tile_cols) *max_tile_cols = widest;
        return nt_;
    };
    int max_cols = g_tc_max_bn < gcols ? gcols : g_tc_max_bn, bn = 0;
    // Resident weights (halo mode, one 64-channel slice, e.g. conv2_3x3 and the 64-channel 3x3 convs of 3a-3c): an N tile
    // of <= 96 columns keeps all nine taps of its weights (<= 108 KB) in shared memory for the life of the CTA, so the
    // only per-tile TMA traffic left is ONE activation halo box.
    // tc_bres = 2: the same with CTA pairs -- each CTA of a pair keeps its half of a <= 192-column weight tile, the MMA
    // keeps the full N (conv2_3x3: 2 x 108 KB resident, N = 192)
    const bool want_pair_res = !grouped && p.halo && g_tc_bres == 2 && p.cchunks == 1 && ksplit == 1 && nseg == 1 &&
                               g_tc_cta2 != 0 && cout > 128 && cout <= 192;
    const bool want_resident = !grouped && p.halo && g_tc_bres == 1 && p.cchunks == 1 && ksplit == 1 && nseg == 1;
    if (want_resident && max_cols > 96) max_cols = 96;
    if (want_pair_res && max_cols > 192) max_cols = 192;
    int nt = partition(max_cols, &bn);
    // underfilled grid: more, narrower N tiles (tiles of <= 128 columns run two CTAs per SM: twice the slots to fill)
    // (a problem of a group whose other members fill the GPU keeps its widest tiles)
    while (!group_fills_gpu &&
           m_tiles * nt * ksplit < (max_cols <= 128 && g_tc_occ2 && g_tc_cta2 != 2 && !want_resident ? 2 * sms : sms) &&
           max_cols > gcols) {
        int w_ = 0;
        const int nt2 = partition(max_cols - 16, &w_);
        if (nt2 > TBN_MAX_NTILES) break;
        max_cols -= 16;
        nt = nt2;
    }
    for (int mc = max_cols - 16; mc >= gcols; mc -= 16) {  // same tile count, evener tiles
        int w_ = 0;
        if (partition(mc, &w_) != nt) break;
        max_cols = mc;
    }
    nt = partition(max_cols, &bn);
    TBN_CHECK_ARG(nt <= TBN_MAX_NTILES, "conv_tc: too many N tiles (%d)", nt);
    p.bn_tile = bn;
    p.n_tiles = nt;
    p.mg_nt = nt == 1 ? 0xFFFFFFFFu : (unsigned)(0x100000000ULL / (unsigned)nt);
    p.mg_tw = p.tiles_w == 1 ? 0xFFFFFFFFu : (unsigned)(0x100000000ULL / (unsigned)p.tiles_w);
    p.mg_th = p.tiles_h == 1 ? 0xFFFFFFFFu : (unsigned)(0x100000000ULL / (unsigned)p.tiles_h);
    p.out_f32 = (out_dtype == TBN_F32); p.bias_period = bias_period;
    p.out_n = ksplit > 1 ? ksplit : p.n;
    p.tmem_cols = bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : 256;
    p.bias = bias;
    p.dbg = g_tc_debug;
    p.trace = g_tc_trace;
    TBN_CHECK_ARG(m_tiles * nt < (1LL << 31), "conv_tc: grid too large");
    TBN_CHECK_ARG(ksplit >= 1 && (ksplit == 1 || (k == 1 && nseg == 1 && out_dtype == TBN_F32)),
                  "conv_tc: split-K needs a GEMM (k = 1) with one fp32 destination");
    TBN_CHECK_ARG(p.cchunks % ksplit == 0, "conv_tc: split-K needs (cin / 64) %% ksplit == 0 (%d k-blocks, ksplit %d)", p.cchunks,
                  ksplit);
    p.ksplit = ksplit;
    p.cc_per_split = cdiv(p.cchunks, ksplit);
    p.mg_ks = ksplit == 1 ? 0xFFFFFFFFu : (unsigned)(0x100000000ULL / (unsigned)ksplit);
    const bool pair_res = want_pair_res && bn <= 192 && bn % 32 == 0;
    p.b_resident = ((want_resident && bn <= 96) || pair_res) ? 1 : 0;
    // CTA pairs: 2 always, 1 for every N tile wider than 128 columns, 3 only where they measured faster than independent
    // CTAs: tap-mode 3x3 convs (long K, one destination) with wide N tiles
    const bool cta2 = pair_res || (ksplit == 1 && !p.b_resident &&
                                   (g_tc_cta2 == 2 || (g_tc_cta2 == 1 && bn > 128) ||
                                    (g_tc_cta2 == 3 && bn > 128 && k == 3 && !p.halo && nseg == 1 && p.cchunks >= 2)));
    p.cta2_mul = cta2 ? 2 : 1;
    p.total_tiles = (int)((cta2 ? (m_tiles + 1) / 2 : m_tiles) * nt) * ksplit;  // CTA-pair mode: tiles = pairs of pixel tiles
    // Narrow tiles (2 accumulators <= 256 TMEM columns): TWO CTAs per SM, each with half the shared memory and one
    // epilogue team.  A single MMA-issuing thread needs ~45 cycles per tcgen05.mma (measured), more than the tensor
    // pipe takes for N <= 96, so two independent pipelines per SM are what keeps the pipe busy on those layers.
    const int ctas_per_sm = (!cta2 && !p.b_resident && 2 * p.tmem_cols <= 256 && g_tc_occ2) ? 2 : 1;
    // Epilogue teams (4 warps each; team t takes column groups t, t + teams, ... of a tile).  Two CTAs per SM: one team
    // each.  Otherwise two, or three when that lowers the widest team's share (N = 160 / 192 / 224: 64-96 columns per
    // team instead of 96-128) -- the short-K 1x1 GEMMs and conv2_3x3 are paced by the epilogue, not by the MMAs.
    p.epi_teams = 1;
    if (ctas_per_sm == 1 && cdiv(bn, gcols) >= 2) {
        int worst[4] = {0, 0, 0, 0};
        for (int teams = 2; teams <= 3; ++teams)
            for (int t = 0; t < nt; ++t) {
                int share[3] = {0, 0, 0};
                for (int g = p.tile_grp[t]; g < p.tile_grp[t + 1]; ++g) share[(g - p.tile_grp[t]) % teams] += p.grp_col[g + 1] - p.grp_col[g];
                for (int i = 0; i < teams; ++i) worst[teams] = share[i] > worst[teams] ? share[i] : worst[teams];
            }
        p.epi_teams = (g_tc_teams3 && worst[3] < worst[2]) ? 3 : 2;
    }

    // shared-memory plan
    p.b_stage_bytes = (unsigned)(cta2 ? bn / 2 : bn) * 128u;  // CTA-pair mode: each CTA stages half of the weight tile
    p.b_tx_bytes = p.b_stage_bytes;
    const int fixed = 1024 /*alignment slack*/ + p.epi_teams * EPI_STAGE_BYTES + 1024 /*barriers*/;
    const int budget = (ctas_per_sm == 2 ? 113 * 1024 : 227 * 1024) - fixed;
    if (p.halo) {
        p.a_tx_bytes = HALO_ROWS * 128u;
        p.a_stage_bytes = (p.a_tx_bytes + 1023u) & ~1023u;
        p.a_stages = (p.cchunks == 1 || ctas_per_sm == 2) ? 2 : 3;
        int bst = (budget - p.a_stages * (int)p.a_stage_bytes) / (int)p.b_stage_bytes;
        if (p.b_resident) {  // exactly one slot per filter tap, filled once
            if (bst >= 9 + (3 - p.a_stages) * 0) {
                p.a_stages = (budget - 9 * (int)p.b_stage_bytes) / (int)p.a_stage_bytes >= 3 ? 3 : 2;
                bst = 9;
            } else {
                p.b_resident = 0;
            }
        }
        if (bst > 16) bst = 16;
        TBN_CHECK_ARG(bst >= 2, "conv_tc: shared memory plan failed (halo)");
        p.b_stages = bst;
    } else {
        p.a_tx_bytes = (unsigned)(p.bw * p.bh * p.bni) * 128u;
        p.a_stage_bytes = A_STAGE_BYTES;
        int st = budget / (int)(p.a_stage_bytes + p.b_stage_bytes);
        if (st > 12) st = 12;
        TBN_CHECK_ARG(st >= 2, "conv_tc: shared memory plan failed");
        p.a_stages = p.b_stages = st;
    }
    {
        const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x) + x_co;
        const int aw = k == 1 ? p.w : w_in, ah = k == 1 ? p.h : h_in;  // input extent (1x1: flattened pixels)
        cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)aw, (cuuint64_t)ah, (cuuint64_t)p.n};
        cuuint64_t strides[3] = {(cuuint64_t)x_cs * 2, (cuuint64_t)x_cs * 2 * aw, (cuuint64_t)x_cs * 2 * aw * ah};
        // stride 2: the box TRAVERSES (b-1)*2+1 pixels with element stride 2, i.e. still loads bw x bh pixels
        cuuint32_t box[4] = {64, (cuuint32_t)((p.bw - 1) * stride + 1), (cuuint32_t)((p.bh - 1) * stride + 1), (cuuint32_t)p.bni};
        if (p.halo) { box[1] = HALO_PITCH; box[2] = HALO_BH + 2; box[3] = 1; }
        cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)xb, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TBN_CHECK_ARG(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        const long long kp = (long long)k * k * p.cin_pad;
        cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)cout};
        cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(cta2 ? bn / 2 : bn)};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TBN_CHECK_ARG(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
    }

    memset(&tmY, 0, sizeof(tmY));
    {
        const int esz = p.out_f32 ? 4 : 2;
        for (int sg = 0; sg < nseg; ++sg) {
            char* yb = reinterpret_cast<char*>(segs[sg].y) + (size_t)segs[sg].y_co * esz;
            const cuuint64_t pitch = (cuuint64_t)segs[sg].y_cs * esz;
            cuuint64_t dims[4] = {(cuuint64_t)segs[sg].cols, (cuuint64_t)p.w, (cuuint64_t)p.h,
                                  (cuuint64_t)(ksplit > 1 ? ksplit : p.n)};
            cuuint64_t strides[3] = {pitch, pitch * p.w, pitch * p.w * p.h};
            cuuint32_t box[4] = {(cuuint32_t)gcols, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bni};
            cuuint32_t es[4] = {1, 1, 1, 1};
            CUresult r = encode(&tmY.m[sg], p.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, yb,
                                dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            TBN_CHECK_ARG(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(Y%d) failed with %d", sg, (int)r);
        }
    }

    pl.cta2 = cta2;
    pl.ctas_per_sm = ctas_per_sm;
    pl.nt = nt;
    pl.bn = bn;
    pl.m_tiles = m_tiles;
    pl.cost = (long long)p.cchunks * k * k * bn;
    return 0;
}

// problems that may share one launch: same kernel variant and the same CTAs-per-SM regime
static bool plans_compatible(const ConvPlan& a, const ConvPlan& b) {
    return a.pr.p.halo == b.pr.p.halo && a.cta2 == b.cta2 && a.ctas_per_sm == b.ctas_per_sm && !a.pr.p.b_resident &&
           !b.pr.p.b_resident && a.pr.p.ksplit == 1 && b.pr.p.ksplit == 1;
}

template <int NP>
static cudaError_t launch_group(const ConvPlan* const* pls, int n, int grid, int threads, size_t smem, bool cta2, bool halo,
                                void* stream) {
    static thread_local TcGroup<NP>* tg = nullptr;  // host staging copy of the kernel parameter (13 KB for NP = 6)
    if (!tg) tg = new TcGroup<NP>;
    TcGroup<NP>& G = *tg;
    G.nprob = n;
    int acc = 0;
    for (int i = 0; i < NP; ++i) {
        if (i < n) {
            G.pr[i] = pls[i]->pr;
            acc += pls[i]->pr.p.total_tiles;
        }
        G.tile_end[i] = acc;
    }
    G.total_tiles = acc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cta2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (g_tc_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    if (cta2) return halo ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, true, NP>, G) : cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, true, NP>, G);
    return halo ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, false, NP>, G) : cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, false, NP>, G);
}

// Launch 1 .. TBN_MAX_PROBS compatible planned problems as ONE grid.  For a group the shared-memory / tensor-memory plan
// is re-derived from the maxima over the problems.
static int launch_plans(ConvPlan** pls, int n, void* stream) {
    TBN_CHECK_ARG(n >= 1 && n <= TBN_MAX_PROBS, "conv_tc: 1..%d problems per launch", TBN_MAX_PROBS);
    const int sms = sm_count();
    const bool cta2 = pls[0]->cta2, halo = pls[0]->pr.p.halo != 0;
    const int ctas_per_sm = pls[0]->ctas_per_sm;
    if (n > 1) {
        // longest tiles first: the static round-robin then ends on the short ones
        for (int i = 1; i < n; ++i)
            for (int j = i; j > 0 && pls[j]->cost > pls[j - 1]->cost; --j) { ConvPlan* t = pls[j]; pls[j] = pls[j - 1]; pls[j - 1] = t; }
        unsigned bslot = 0, tmem = 0;
        int teams = 1, a_stages_halo = 2;
        for (int i = 0; i < n; ++i) {
            const TcParams& q = pls[i]->pr.p;
            TBN_CHECK_ARG(plans_compatible(*pls[0], *pls[i]), "conv_tc: incompatible problems in one group");
            bslot = q.b_tx_bytes > bslot ? q.b_tx_bytes : bslot;
            tmem = q.tmem_cols > tmem ? q.tmem_cols : tmem;
            teams = q.epi_teams > teams ? q.epi_teams : teams;
            if (q.halo && q.cchunks > 1 && ctas_per_sm == 1) a_stages_halo = 3;
        }
        const int fixed = 1024 + teams * EPI_STAGE_BYTES + 1024;
        const int budget = (ctas_per_sm == 2 ? 113 * 1024 : 227 * 1024) - fixed;
        int a_stages, b_stages;
        unsigned a_stage_bytes;
        if (halo) {
            a_stage_bytes = ((unsigned)HALO_ROWS * 128u + 1023u) & ~1023u;
            a_stages = a_stages_halo;
            b_stages = (budget - a_stages * (int)a_stage_bytes) / (int)bslot;
            if (b_stages > 16) b_stages = 16;
            TBN_CHECK_ARG(b_stages >= 2, "conv_tc: shared memory plan failed (halo group)");
        } else {
            a_stage_bytes = A_STAGE_BYTES;
            int st = budget / (int)(a_stage_bytes + bslot);
            if (st > 12) st = 12;
            TBN_CHECK_ARG(st >= 2, "conv_tc: shared memory plan failed (group)");
            a_stages = b_stages = st;
        }
        for (int i = 0; i < n; ++i) {
            TcParams& q = pls[i]->pr.p;
            q.b_stage_bytes = bslot; q.tmem_cols = tmem; q.epi_teams = teams;
            q.a_stages = a_stages; q.b_stages = b_stages; q.a_stage_bytes = a_stage_bytes;
        }
    }
    const TcParams& p = pls[0]->pr.p;
    const size_t smem = 1024 + (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * p.b_stage_bytes +
                        (size_t)p.epi_teams * EPI_STAGE_BYTES + 8 * (2 * p.a_stages + 2 * p.b_stages + 4) + 16;
    TBN_CHECK_ARG(smem <= 227 * 1024, "conv_tc: shared memory plan exceeds 227 KB");
    long long total = 0;
    for (int i = 0; i < n; ++i) total += pls[i]->pr.p.total_tiles;
    TBN_CHECK_ARG(total < (1LL << 31), "conv_tc: grid too large");

    static DeviceOnce once;
    int once_dev = 0;
    if (once.needed(&once_dev)) {
        cudaError_t e = cudaSuccess;
        const void* fns[8] = {(const void*)conv_tc_kernel<false, false, 1>, (const void*)conv_tc_kernel<true, false, 1>,
                              (const void*)conv_tc_kernel<false, true, 1>, (const void*)conv_tc_kernel<true, true, 1>,
                              (const void*)conv_tc_kernel<false, false, TBN_MAX_PROBS>, (const void*)conv_tc_kernel<true, false, TBN_MAX_PROBS>,
                              (const void*)conv_tc_kernel<false, true, TBN_MAX_PROBS>, (const void*)conv_tc_kernel<true, true, TBN_MAX_PROBS>};
        for (int i = 0; i < 8 && e == cudaSuccess; ++i)
            e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "conv_tc smem opt-in");
        once.done(once_dev);
    }
    const int threads = 64 + 128 * p.epi_teams;
    int grid;
    if (cta2) {
        int clusters = total < sms / 2 ? (int)total : sms / 2;
        if (n == 1 && pls[0]->nt > 1 && clusters >= pls[0]->nt) clusters -= clusters % pls[0]->nt;
        grid = 2 * clusters;
    } else {
        grid = total < (long long)sms * ctas_per_sm ? (int)total : sms * ctas_per_sm;
        // a CTA then keeps one N tile: its weight boxes stay hot in L2
        if (n == 1 && pls[0]->nt > 1 && grid >= pls[0]->nt) grid -= grid % pls[0]->nt;
    }
    const ConvPlan* const* cp = pls;
    cudaError_t e = n == 1 ? launch_group<1>(cp, n, grid, threads, smem, cta2, halo, stream)
                           : launch_group<TBN_MAX_PROBS>(cp, n, grid, threads, smem, cta2, halo, stream);
    if (e != cudaSuccess) return cuda_fail(e, "conv_tc launch");
    TBN_LAUNCH_CHECK("conv_tc launch");
    return 0;
}

static int launch_conv_tc(const void* x, const void* wp, const float* bias, const OutSeg* segs, int nseg, int n, int h_in,
                          int w_in, int cin, int x_cs, int x_co, int k, int stride, int out_dtype, int bias_period,
                          void* stream, int ksplit = 1) {
    static thread_local ConvPlan* pl = nullptr;
    if (!pl) pl = new ConvPlan;
    int rc = plan_conv_tc(*pl, x, wp, bias, segs, nseg, n, h_in, w_in, cin, x_cs, x_co, k, stride, out_dtype, bias_period, ksplit,
                          false);
    if (rc) return rc;
    ConvPlan* one[1] = {pl};
    return launch_plans(one, 1, stream);
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_set_tuning(const char* key, int value) {
    TBN_CHECK_ARG(key != nullptr, "set_tuning: null key");
    if (!strcmp(key, "tc_halo")) g_tc_halo = value < 0 ? 0 : (value > 2 ? 2 : value);
    else if (!strcmp(key, "tc_debug")) g_tc_debug = value;
    else if (!strcmp(key, "tc_bres")) g_tc_bres = value < 0 ? 0 : (value > 2 ? 2 : value);
    else if (!strcmp(key, "tc_pdl")) g_tc_pdl = value ? 1 : 0;
    else if (!strcmp(key, "tc_max_bn")) g_tc_max_bn = value < 32 ? 32 : (value > 256 ? 256 : value / 16 * 16);
    else if (!strcmp(key, "tc_occ2")) g_tc_occ2 = value ? 1 : 0;
    else if (!strcmp(key, "tc_teams3")) g_tc_teams3 = value ? 1 : 0;
    else if (!strcmp(key, "tc_cta2")) g_tc_cta2 = value;
    else TBN_CHECK_ARG(false, "set_tuning: unknown key '%s'", key);
    return 0;
}

extern "C" int tbn_debug_trace(void* dev_buf, int* ctas, int* stride, int* n) {
    g_tc_trace = reinterpret_cast<unsigned long long*>(dev_buf);
    if (ctas) *ctas = TRACE_CTAS;
    if (stride) *stride = TRACE_STRIDE;
    if (n) *n = TRACE_N;
    return 0;
}

extern "C" int tbn_conv_tc(const void* x, const void* wp, const float* bias, void* y, int n, int h, int w, int cin,
                           int x_cs, int x_co, int cout, int y_cs, int y_co, int k, int relu, int out_dtype,
                           int bias_period, void* stream) {
    OutSeg seg{y, y_cs, y_co, cout, TBN_SEG_BIAS | (relu ? TBN_SEG_RELU : 0)};
    return launch_conv_tc(x, wp, bias, &seg, 1, n, h, w, cin, x_cs, x_co, k, 1, out_dtype, bias_period, stream);
}

extern "C" int tbn_conv_tc_group(const tbn_conv_desc* descs, int n, void* stream) {
    TBN_CHECK_ARG(descs != nullptr && n >= 1 && n <= 16, "conv_tc_group: 1..16 problems");
    static thread_local ConvPlan* plans = nullptr;
    if (!plans) plans = new ConvPlan[16];
    for (int pass = 0; pass < 2; ++pass) {
        long long tiles = 0;
        for (int i = 0; i < n; ++i) {
            const tbn_conv_desc& d = descs[i];
            TBN_CHECK_ARG(d.segs != nullptr && d.nseg >= 1 && d.nseg <= TBN_MAX_SEGS, "conv_tc_group: bad segment list (problem %d)", i);
            OutSeg segs[TBN_MAX_SEGS];
            for (int s = 0; s < d.nseg; ++s) segs[s] = OutSeg{d.segs[s].y, d.segs[s].y_cs, d.segs[s].y_co, d.segs[s].cols, d.segs[s].flags};
            // pass 0: every problem keeps its widest N tiles (the group as a whole is expected to fill the GPU)
            int rc = plan_conv_tc(plans[i], d.x, d.wp, d.bias, segs, d.nseg, d.n, d.h, d.w, d.cin, d.x_cs, d.x_co, d.k, d.stride,
                                  d.out_dtype, 1, 1, n > 1, n > 1 && pass == 0);
            if (rc) return rc;
            tiles += plans[i].pr.p.total_tiles;
        }
        if (n == 1 || pass == 1 || tiles >= sm_count()) break;  // too few tiles even together: plan each for itself
    }
    // greedy partition into launches of compatible problems (same kernel variant), in the caller's order
    bool done[16] = {false};
    for (int i = 0; i < n; ++i) {
        if (done[i]) continue;
        ConvPlan* grp[TBN_MAX_PROBS];
        int m = 0;
        grp[m++] = &plans[i];
        done[i] = true;
        for (int j = i + 1; j < n && m < TBN_MAX_PROBS; ++j)
            if (!done[j] && plans_compatible(plans[i], plans[j])) {
                grp[m++] = &plans[j];
                done[j] = true;
            }
        int rc = launch_plans(grp, m, stream);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int tbn_gemm_tc_splitk_ld(const void* x, int ldx, const void* wp, void* partial, int m, int k, int n, int ksplit,
                                     void* stream) {
    static const float zero_bias[4] = {0, 0, 0, 0};  // never read: the segment carries no TBN_SEG_BIAS flag
    TBN_CHECK_ARG(ldx >= k && ldx % 8 == 0, "gemm_tc_splitk: row stride must be >= k and a multiple of 8");
    OutSeg seg{partial, n, 0, n, 0};
    return launch_conv_tc(x, wp, zero_bias, &seg, 1, 1, 1, m, k, ldx, 0, 1, 1, TBN_F32, 1, stream, ksplit);
}

extern "C" int tbn_gemm_tc_splitk(const void* x, const void* wp, void* partial, int m, int k, int n, int ksplit, void* stream) {
    return tbn_gemm_tc_splitk_ld(x, k, wp, partial, m, k, n, ksplit, stream);
}

extern "C" int tbn_conv_tc_s2(const void* x, const void* wp, const float* bias, void* y, int n, int h, int w, int cin,
                              int x_cs, int x_co, int cout, int y_cs, int y_co, int relu, int out_dtype, void* stream) {
    OutSeg seg{y, y_cs, y_co, cout, TBN_SEG_BIAS | (relu ? TBN_SEG_RELU : 0)};
    return launch_conv_tc(x, wp, bias, &seg, 1, n, h, w, cin, x_cs, x_co, 3, 2, out_dtype, 1, stream);
}

extern "C" int tbn_conv_tc_multi(const void* x, const void* wp, const float* bias, const tbn_out_seg* segs_host,
                                 int nseg, int n, int h, int w, int cin, int x_cs, int x_co, int k, int out_dtype,
                                 void* stream) {
    TBN_CHECK_ARG(segs_host != nullptr && nseg >= 1 && nseg <= TBN_MAX_SEGS, "conv_tc_multi: bad segment list");
    OutSeg segs[TBN_MAX_SEGS];
    for (int s = 0; s < nseg; ++s)
        segs[s] = OutSeg{segs_host[s].y, segs_host[s].y_cs, segs_host[s].y_co, segs_host[s].cols, segs_host[s].flags};
    return launch_conv_tc(x, wp, bias, segs, nseg, n, h, w, cin, x_cs, x_co, k, 1, out_dtype, 1, stream);
}
