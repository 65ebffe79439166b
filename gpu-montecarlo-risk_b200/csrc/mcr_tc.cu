// The fused hot-path kernel, tensor-core flavour: batched correlate  D[128 x N] = Z[128 x K] * B^T
// on the 5th-generation tensor cores (tcgen05.mma, accumulator in TMEM), normals staged in shared
// memory only, threshold/count epilogue straight out of TMEM.
//
// Same pipeline as the FFMA kernel (mcr_ffma.cuh) and the reference trial
// (src/02-C-serial/monte_carlo_serial.c:62-85), restructured as a real GEMM over a tile of 128 trials:
//   generator warps : Philox4x32-10 -> Box-Muller normals z (registers) -> split z = z_hi + z_lo (fp16)
//                     -> written K-major, 128-byte-swizzled into shared memory (never to HBM)
//   MMA warp        : one elected thread issues  D  = Zhi*Bhi^T + Zhi*Blo^T + Zlo*Bhi^T   (fp16x3,
//                     fp32 accumulate in TMEM); B = the transform, split the same way, resident in smem.
//                     Column K-1.. holds the constant term: z[:, N] = 1 and B[i][N] = mu_i + x, so the
//                     accumulator is directly  R_i + x  (padded rows get +1 and can never hit).
//   epilogue        : the generator warps read their 32 TMEM lanes (tcgen05.ld), count sign bits,
//                     track the guard band, and reduce like the FFMA kernel; near-threshold trials go
//                     to the same fp64 fix-up list.
// Why fp16x3 and not TF32x3: the split operands are half the size (Z tiles + B stay resident for
// N <= 128) and hi + lo carries 22 significand bits, so the dropped lo*lo term is 2^-22 relative:
// the guard band is about as narrow as the fp32 FFMA kernel's and the final counts remain
// bit-identical to fp64 on the same normals.  (bf16x3 was tried first: 2^-16, four times as many
// trials sent to the fp64 fix-up.)
#include "../../include/mcr_b200.h"
#include "mcr_internal.h"
#include "mcr_rng.cuh"
#include "mcr_tc_common.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mcr {


constexpr uint32_t PORT_BINS = MCR_PORTFOLIO_BINS;

struct TcParams {
    uint64_t M, path_offset;
    uint32_t key0, key1;
    int32_t k, N;
    float delta_rel, delta_abs;   // guard band = delta_rel * max_j |z_j| + delta_abs (0, 0 disables the recheck)
    const uint4* B_image;   // global: [Bhi image | Blo image], each KB * NOUT * 128 bytes
    int32_t* S_out;
    uint32_t* flag_list;
    uint32_t* flag_count;
    uint32_t flag_cap;
    unsigned long long* pair_list;   // single near-threshold decisions (trial << 16 | S_rest << 8 | row), or nullptr
    uint32_t* pair_count;
    uint32_t pair_cap;
    uint64_t* block_partials;
    // portfolio extension (north-star A7; no reference code): accumulator column NOUT-1 holds w.R
    int32_t portfolio;            // 0 = off (column NOUT-1 is an ordinary padded / asset row)
    float port_thresh;            // scaled: tail event is  D < port_thresh  (= -x_p * TC_SCALE)
    float hist_lo, hist_inv_w;    // unscaled histogram origin and 1 / bin width
    double* port_acc;             // [4] += hits, sum P, sum P^2, sum of -P over tail events
    unsigned long long* port_hist;  // [MCR_PORTFOLIO_BINS] counts
    double* port_hist_sum;          // [MCR_PORTFOLIO_BINS] per-bin sum of P
    // the 8-column chunk holding column N: z <- z * tail_mul + tail_add  (normals | the constant 1 | zeros)
    float tail_mul[8], tail_add[8];
};


#ifndef MCR_TC_GROUPS
#define MCR_TC_GROUPS 6   // warpgroups per CTA where TMEM and shared memory allow (768 threads: 80 registers each, no spills)
#endif
#ifndef MCR_TC_UNROLL
#define MCR_TC_UNROLL 2
#endif
constexpr int TC_UNROLL = MCR_TC_UNROLL;  // 8-normal chunks in flight per generator thread


// NOUT   = transform rows padded to a multiple of 16 (UMMA N)
// KB     = 64-column K-blocks (K = N + 1 padded)
// GROUPS = warpgroups per CTA (one CTA per SM; the groups are what fills the SM).
//
// Every warpgroup (128 threads = the 128 trials of a tile) is a self-contained pipeline; the groups share
// only the resident transform image:
//   * the group owns ONE shared-memory slot (a K-block of Z: 128 trials x 64 columns, fp16 hi + lo =
//     32 KB), ONE TMEM accumulator region and ONE mbarrier.  K-block u of the group is generated into the
//     slot, handed over through a named barrier of the group's 128 threads, and one warp of the group
//     (warp u % 4, so the extra work rotates) issues its MMAs and commits them to the mbarrier: phase u
//     of the mbarrier completes when the tensor core has finished K-blocks 0..u.
//   * before the first store of K-block u + 1 the threads wait for phase u (the slot is free again and,
//     when u closed a tile, the tile's accumulator is complete).  The wait comes after the first batch of
//     Philox / Box-Muller work of that K-block, and the MMAs of a K-block take a few hundred cycles, so
//     it does not stall in steady state.
//   * the epilogue of tile n runs after K-block 0 of tile n + 1 has been generated and before that
//     K-block's hand-over barrier: everyone has read the accumulator region before the MMA that
//     overwrites it is issued, so a single region per group is enough.
// No group ever waits for another one (an earlier version served all groups from one MMA warp in
// round-robin order: about 10 % of the warp samples sat in that coupling).
template <int NOUT, int KB, int GROUPS>
__global__ void __launch_bounds__(128 * GROUPS, 1) tc_path_kernel(const TcParams p) {
    using namespace tc;
    constexpr uint32_t SLOT_BYTES = 2 * TILE_M * 128;   // one K-block of Z: hi then lo
    constexpr uint32_t B_BYTES = KB * NOUT * 128;       // one of Bhi / Blo
    constexpr uint32_t GROUP_COLS = (NOUT + 31) & ~31;  // TMEM columns of one accumulator region
    constexpr uint32_t NEED_COLS = GROUP_COLS * GROUPS;
    constexpr uint32_t TMEM_COLS = NEED_COLS <= 32 ? 32 : NEED_COLS <= 64 ? 64 : NEED_COLS <= 128 ? 128 : NEED_COLS <= 256 ? 256 : 512;
    static_assert(NEED_COLS <= 512, "accumulators do not fit TMEM");
    constexpr int NTHREADS = 128 * GROUPS;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* s_ring = smem;                                   // [group][hi|lo][128 x 128 B]
    uint8_t* s_bhi = smem + GROUPS * SLOT_BYTES;
    uint8_t* s_blo = s_bhi + B_BYTES;
    uint32_t* s_hist_cnt = reinterpret_cast<uint32_t*>(s_blo + B_BYTES);          // only when p.portfolio (host sizes smem)
    float* s_hist_sum = reinterpret_cast<float*>(s_hist_cnt + PORT_BINS);
    __shared__ uint64_t s_mma[GROUPS];
    __shared__ double s_port[4 * GROUPS][4];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_cnt[4 * GROUPS], s_sum[4 * GROUPS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup ----
    {
        const uint4* src = p.B_image;
        uint4* dst = reinterpret_cast<uint4*>(s_bhi);
        for (uint32_t t = tid; t < 2 * B_BYTES / 16; t += NTHREADS) dst[t] = src[t];
    }
    if (p.portfolio)
        for (uint32_t t = tid; t < PORT_BINS; t += NTHREADS) { s_hist_cnt[t] = 0u; s_hist_sum[t] = 0.0f; }
    if (tid == 0)
        for (int gg = 0; gg < GROUPS; ++gg) mbar_init(&s_mma[gg], 1);
    if (warp == 0) tmem_alloc(&s_tmem, TMEM_COLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    const uint64_t n_tiles = (p.M + TILE_M - 1) / TILE_M;
    // tiles of this CTA: global tile = blockIdx.x + gridDim.x * m, m = 0, 1, ...; group = m % GROUPS
    const uint64_t cta_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    constexpr uint32_t IDESC = make_idesc(NOUT);
    uint32_t cnt = 0, sum_s = 0;
    uint32_t p_hits = 0;
    float p_sum = 0.f, p_sum2 = 0.f, p_tail = 0.f;

    {
        const int g = warp / 4, wg = warp % 4;
        const int row = tid - 128 * g;
        uint64_t* const bar = &s_mma[g];
        const uint32_t slot_base = smem_u32(s_ring + g * SLOT_BYTES);
        const uint32_t zhi_row = slot_base + (uint32_t)(row / 8) * 1024u + (uint32_t)(row % 8) * 128u, zlo_row = zhi_row + TILE_M * 128;
        const uint32_t r7s = (uint32_t)(row % 8) << 4;
        const uint32_t d_tmem = tmem_base + g * GROUP_COLS;
        const uint32_t taddr = d_tmem + ((uint32_t)(wg * 32) << 16);
        const uint32_t bhi = smem_u32(s_bhi), blo = smem_u32(s_blo);
        const uint64_t my_tiles = cta_tiles > (uint64_t)g ? (cta_tiles - g + GROUPS - 1) / GROUPS : 0;
        const int n_full = p.N / 8;       // chunks (8 columns) made of normals only
        float prev_zmax = 0.0f;
        uint64_t prev_i = 0;
        uint32_t u = 0;                   // K-blocks this group has produced so far

        // epilogue of a finished tile: TMEM -> registers, threshold + guard band
        auto epilogue = [&](uint64_t i, float zmax) {
            tc_fence_after();
            // Near-threshold decisions are rare.  The scan below counts how many of the trial's decisions are inside
            // its guard band and remembers the last one (column, fast-path decision); it runs only where the smallest
            // |accumulator| is under the band: per 32-column group (wide tiles), or after the fast pass on a second
            // read of the accumulators when any lane of the warp needs it (narrow tiles, where that is cheaper).
            constexpr bool SCAN_INLINE = NOUT > 64;
            const float thr = __fmaf_rn(p.delta_rel, zmax, p.delta_abs);
            uint32_t S = 0, n_amb = 0, amb = 0;
            float band = 3.0e38f;
            auto scan = [&](const uint32_t (&v)[32], int c0) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (c0 + c < NOUT - 1 || (c0 + c == NOUT - 1 && !p.portfolio)) {
                        if (fabsf(__uint_as_float(v[c])) < thr) { ++n_amb; amb = ((uint32_t)(c0 + c) << 1) | (v[c] >> 31); }
                    }
                }
            };
#pragma unroll
            for (int c0 = 0; c0 < (MCR_TC_ABLATE == 2 ? 0 : NOUT); c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                tmem_ld_wait();
                if (SCAN_INLINE) band = 3.0e38f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (c0 + c < NOUT - 1) {
                        S += v[c] >> 31;
                        band = fminf(band, fabsf(__uint_as_float(v[c])));
                    } else if (c0 + c == NOUT - 1) {
                        if (!p.portfolio) {
                            S += v[c] >> 31;
                            band = fminf(band, fabsf(__uint_as_float(v[c])));
                        } else if (i < p.M) {
                            // portfolio return of this trial (the host put w^T A and w.mu in row NOUT-1)
                            const float Ps = __uint_as_float(v[c]);
                            const float P = Ps * (float)(1.0 / TC_SCALE);
                            const bool tail = Ps < p.port_thresh;
                            p_hits += tail ? 1u : 0u;
                            p_tail += tail ? -P : 0.0f;
                            p_sum += P;
                            p_sum2 = fmaf(P, P, p_sum2);
                            int bin = (int)floorf((P - p.hist_lo) * p.hist_inv_w);
                            bin = bin < 0 ? 0 : (bin >= (int)PORT_BINS ? (int)PORT_BINS - 1 : bin);
                            atomicAdd(&s_hist_cnt[bin], 1u);
                            atomicAdd(&s_hist_sum[bin], P);
                        }
                    }
                }
                if (SCAN_INLINE && band < thr) scan(v, c0);
            }
            if (!SCAN_INLINE && __any_sync(0xffffffffu, band < thr)) {   // a few warps in a hundred: read the accumulators again
#pragma unroll 1
                for (int c0 = 0; c0 < NOUT; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    tmem_ld_wait();
                    scan(v, c0);
                }
            }
            if (i < p.M) {
                if (n_amb == 1 && p.pair_list) {
                    // one uncertain decision: exact_pairs_kernel re-evaluates that row in fp64 and counts the trial
                    const uint32_t slot = atomicAdd(p.pair_count, 1u);
                    if (slot < p.pair_cap)
                        p.pair_list[slot] = ((unsigned long long)i << 16) | (unsigned long long)(((S - (amb & 1u)) << 8) | (amb >> 1));
                } else if (n_amb != 0) {
                    const uint32_t slot = atomicAdd(p.flag_count, 1u);
                    if (slot < p.flag_cap) p.flag_list[slot] = (uint32_t)i;
                } else {
                    cnt += (S >= (uint32_t)p.k) ? 1u : 0u;
                    sum_s += S;
                    if (p.S_out) p.S_out[i] = (int32_t)S;
                }
            }
        };

        for (uint64_t n = 0; n < my_tiles; ++n) {
            const uint64_t tile = blockIdx.x + (uint64_t)gridDim.x * (n * GROUPS + g);
            const uint64_t i = tile * TILE_M + row;   // chunk-relative trial index of this thread
            const uint64_t path = p.path_offset + i;
            PhiloxPath rng;
            rng.init((uint32_t)path, (uint32_t)(path >> 32), p.key1);
            float zmax = 0.0f;
#pragma unroll 1
            for (int kb = 0; kb < KB; ++kb, ++u) {
                // the slot may be overwritten once the MMAs of the previous K-block have completed (phase u - 1)
                bool slot_free = (u == 0);
                auto acquire = [&]() {
                    // single-K-block tiles: a 100 ns back-off between polls measured 1 % faster; with more K-blocks it did not
                    if (!slot_free) { mbar_wait<(KB == 1 ? 100 : 0)>(bar, (u - 1) & 1u); slot_free = true; }
                };
                // chunks 8*kb .. 8*kb+7 of the row: normals, then the constant 1 in column N, zeros beyond
                const int c_begin = 8 * kb;
                int c8 = c_begin;
                for (; c8 + TC_UNROLL <= min(n_full, c_begin + 8); c8 += TC_UNROLL) {
                    float z[8 * TC_UNROLL];
                    rng.normals_multi<2 * TC_UNROLL>((uint32_t)(2 * c8), p.key0, p.key1, z);
                    acquire();
#pragma unroll
                    for (int w = 0; w < TC_UNROLL; ++w) {
                        float zc[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) zc[q] = z[8 * w + q];
                        store_chunk(zc, c8 + w - c_begin, zhi_row, zlo_row, r7s, zmax);
                    }
                }
                for (; c8 < min(n_full, c_begin + 8); ++c8) {
                    float z[8];
                    rng.normals_multi<2>((uint32_t)(2 * c8), p.key0, p.key1, z);
                    acquire();
                    store_chunk(z, c8 - c_begin, zhi_row, zlo_row, r7s, zmax);
                }
                acquire();
                for (; c8 < c_begin + 8; ++c8) {   // at most 8 such chunks per tile: the one holding column N, then zeros
                    float z[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) z[q] = 0.0f;
                    float zmax_unused = 0.0f;
                    if (c8 == n_full) {
                        if (p.N - 8 * c8 > 4) rng.normals_multi<2>((uint32_t)(2 * c8), p.key0, p.key1, z);
                        else if (p.N - 8 * c8 > 0) rng.normals((uint32_t)(2 * c8), p.key0, p.key1, z);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float zq = z[q] * p.tail_mul[q];        // normals of columns < N, zeros otherwise
                            zmax = fmaxf(zmax, fabsf(zq));
                            z[q] = zq + p.tail_add[q];                    // the constant 1 in column N
                        }
                    }
                    store_chunk(z, c8 - c_begin, zhi_row, zlo_row, r7s, zmax_unused);
                }
                // phase u - 1 is complete here; when it closed tile n - 1, that tile's accumulator is final
                if (kb == 0 && n > 0) epilogue(prev_i, prev_zmax);
                fence_async_smem();     // generic-proxy smem writes -> visible to the tensor core
                tc_fence_before();      // our TMEM reads are ordered before the hand-over
                if (wg == (int)(u & 3u)) {
                    named_bar_sync(1 + g, 128);   // the K-block is complete and everyone is done with the accumulator
                    if (lane == 0) {
                        tc_fence_after();
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
                            const uint32_t za = pass == 2 ? slot_base + TILE_M * 128 : slot_base;
                            const uint32_t ba = (pass == 1 ? blo : bhi) + kb * (NOUT * 128);
#pragma unroll
                            for (int ks = 0; ks < (MCR_TC_ABLATE == 3 ? 0 : KBLOCK / 16); ++ks)
                                umma_f16(d_tmem, make_desc(za + ks * 32), make_desc(ba + ks * 32), IDESC,
                                         (kb | pass | ks) != 0 ? 1u : 0u);
                        }
                        umma_commit(bar);
                    }
                    __syncwarp();
                } else {
                    named_bar_arrive(1 + g, 128);
                }
            }
            prev_i = i;
            prev_zmax = zmax;
        }
        if (my_tiles > 0) {
            mbar_wait(bar, (u - 1) & 1u);
            epilogue(prev_i, prev_zmax);
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        sum_s = __reduce_add_sync(0xffffffffu, sum_s);
        if (lane == 0) { s_cnt[warp] = cnt; s_sum[warp] = sum_s; }
        if (p.portfolio) {
            double d[4] = {(double)p_hits, (double)p_sum, (double)p_sum2, (double)p_tail};
#pragma unroll
            for (int q = 0; q < 4; ++q)
                for (int o = 16; o > 0; o >>= 1) d[q] += __shfl_down_sync(0xffffffffu, d[q], o);
            if (lane == 0)
                for (int q = 0; q < 4; ++q) s_port[warp][q] = d[q];
        }
    }

    // ---- block reduction + teardown ----
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        uint64_t c = 0, s = 0;
        for (int w = 0; w < 4 * GROUPS; ++w) { c += s_cnt[w]; s += s_sum[w]; }
        p.block_partials[2 * (uint64_t)blockIdx.x] = c;
        p.block_partials[2 * (uint64_t)blockIdx.x + 1] = s;
    }
    if (p.portfolio) {
        if (tid < 4) {
            double a = 0;
            for (int w = 0; w < 4 * GROUPS; ++w) a += s_port[w][tid];
            atomicAdd(&p.port_acc[tid], a);
        }
        for (uint32_t t = tid; t < PORT_BINS; t += NTHREADS) {
            const uint32_t c = s_hist_cnt[t];
            if (c) {
                atomicAdd(&p.port_hist[t], (unsigned long long)c);
                atomicAdd(&p.port_hist_sum[t], (double)s_hist_sum[t]);
            }
        }
    }
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ======================================================================================================
// Large asset counts (N > 128): row slabs.
//
// 128 trials x N accumulators do not fit TMEM (512 columns) and the transform does not fit shared memory,
// so the correlate is done one slab of 256 transform rows at a time: slab t needs the normals of
// columns [0, ncols_t) only (ncols_t = min(N, 256 (t+1)) for a lower-triangular transform), which the
// generator simply re-derives from the counter-based stream, and its 256 x 64 transform K-blocks are
// streamed from L2 into the same ring slot as the Z K-block they multiply.  Each launch adds the slab's
// hit count into a per-trial uint16 (bit 15 = "inside the guard band in some slab"); slab_count_kernel
// then reduces that array exactly like the single-launch kernels do.  Costs: the normals are generated
// about (slabs + 1) / 2 times, and 4 bytes of HBM traffic per trial and slab.
// ======================================================================================================
struct TcSlabParams {
    uint64_t M, path_offset;
    uint32_t key0, key1;
    int32_t ncols;        // normal columns this slab multiplies; the constant 1 sits in column ncols
    int32_t kb_count;     // K-blocks = ceil((ncols + 1) / 64)
    float delta_rel, delta_abs;   // scaled guard band = delta_rel * ||z||_2 + delta_abs (Cauchy-Schwarz bound on sum |a||z|)
    float thr_max;                // upper bound of that band over all trials: every |z_j| < 7 (Box-Muller on 32-bit uniforms)
    const uint4* B_slab;  // global: kb_count x [hi 32 KB | lo 32 KB], SLAB_ROWS rows x 128 B, pre-swizzled
    uint16_t* S_acc;      // [M] running hit count, bit 15 = flagged
    int32_t first_slab;   // overwrite S_acc instead of accumulating
    int32_t row_base;     // first transform row of this slab
    unsigned long long* pair_list;   // (trial << 16 | row) of single near-threshold decisions, fixed up per pair
    uint32_t* pair_count;
    uint32_t pair_cap;
};

constexpr int SLAB_ROWS = 256;                            // transform rows per slab = UMMA N = TMEM columns of one accumulator region
constexpr uint32_t SLAB_Z_BYTES = 2 * tc::TILE_M * 128;    // Z K-block: hi + lo
constexpr uint32_t SLAB_B_BYTES = 2 * SLAB_ROWS * 128;     // B K-block: hi + lo
constexpr uint32_t SLAB_SLOT_BYTES = SLAB_Z_BYTES + SLAB_B_BYTES;
constexpr int SLAB_KSPLIT = 4;                             // generator threads per trial row (each owns 16 of a K-block's 64 columns)
constexpr int SLAB_GEN_THREADS = 128 * SLAB_KSPLIT;
constexpr int SLAB_THREADS = SLAB_GEN_THREADS + 32;
constexpr uint32_t SLAB_KEY_MASK = 0xFFFFFC00u;            // epilogue keys: bits 10.. = |accumulator| << 1 (truncated)

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory"); }

// 16 generator warps: warp w works on trial rows 32 (w % 4) .. +31 (its TMEM lane quadrant) and on columns
// 16 (w / 4) .. +15 of every K-block while generating, accumulator columns 64 (w / 4) .. +63 in the epilogue.
__global__ void __launch_bounds__(SLAB_THREADS, 1) tc_slab_kernel(const TcSlabParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t s_empty[2], s_done[2];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_part_S[2][SLAB_KSPLIT][128], s_part_k1[2][SLAB_KSPLIT][128], s_part_k2[2][SLAB_KSPLIT][128];
    __shared__ float s_part_zmax[2][SLAB_KSPLIT][128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t TMEM_COLS = 2 * SLAB_ROWS;   // two accumulator regions of SLAB_ROWS columns: all of TMEM
    constexpr int MMA_WARP = SLAB_GEN_THREADS / 32;
    constexpr int EPI_COLS = SLAB_ROWS / SLAB_KSPLIT;   // accumulator columns per epilogue thread

    if (tid == 0)
        for (int b = 0; b < 2; ++b) { mbar_init(&s_empty[b], 1); mbar_init(&s_done[b], 1); }
    if (warp == MMA_WARP) tmem_alloc(&s_tmem, TMEM_COLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    const uint64_t n_tiles = (p.M + TILE_M - 1) / TILE_M;
    const uint64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    constexpr uint32_t IDESC = make_idesc(SLAB_ROWS);
    const int KBN = p.kb_count;

    if (warp == MMA_WARP) {
        for (uint64_t n = 0; n < my_tiles; ++n) {
            for (int kb = 0; kb < KBN; ++kb) {
                const uint32_t u = (uint32_t)(n * KBN + kb), slot = u & 1;
                named_bar_sync(1 + slot, SLAB_THREADS);
                if (lane == 0) {
                    tc_fence_after();
                    const uint32_t zhi = smem_u32(smem + slot * SLAB_SLOT_BYTES), zlo = zhi + TILE_M * 128;
                    const uint32_t bhi = zhi + SLAB_Z_BYTES, blo = bhi + SLAB_ROWS * 128;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(n & 1) * SLAB_ROWS;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t za = pass == 2 ? zlo : zhi;
                        const uint32_t ba = pass == 1 ? blo : bhi;
#pragma unroll
                        for (int ks = 0; ks < KBLOCK / 16; ++ks)
                            umma_f16(d_tmem, make_desc(za + ks * 32), make_desc(ba + ks * 32), IDESC, (kb | pass | ks) != 0 ? 1u : 0u);
                    }
                    umma_commit(&s_empty[slot]);
                    if (kb == KBN - 1) umma_commit(&s_done[n & 1]);
                }
                __syncwarp();
            }
        }
    } else {
        const int row = tid & 127, part = tid >> 7;   // part: K-split index while generating, column quarter in the epilogue
        const uint32_t ring_row = smem_u32(smem) + (uint32_t)(row / 8) * 1024u + (uint32_t)(row % 8) * 128u;
        const uint32_t r7s = (uint32_t)(row % 8) << 4;
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)part * EPI_COLS;
        const int n_full = p.ncols / 8;
        uint32_t empty_phase = 0, done_phase = 0;
        float prev_zmax = 0.0f;
        uint64_t prev_i = 0;

        auto epilogue = [&](uint64_t n, uint64_t i, float zmax) {
            const uint32_t r = (uint32_t)(n & 1);
            mbar_wait(&s_done[r], (done_phase >> r) & 1u);
            done_phase ^= 1u << r;
            tc_fence_after();
            // Per accumulator a sortable key: |v| (14 mantissa bits) in the high bits, the column in bits 1..8, the sign
            // in bit 0 (rotate left by one, then overwrite the low mantissa bits).  The two smallest keys of the
            // trial's row tell whether none, exactly one (which one, and how the fast path decided it) or several
            // of its decisions are inside the guard band.
            uint32_t S = 0, k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
#pragma unroll
            for (int c0 = 0; c0 < EPI_COLS; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + r * SLAB_ROWS + c0, v);
                tmem_ld_wait();
                float m = 3.0e38f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    S += v[c] >> 31;
                    m = fminf(m, fabsf(__uint_as_float(v[c])));
                }
                // keys only where something can be inside the band (thr_max bounds every trial's band from above)
                if (m < p.thr_max) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const uint32_t key = (__funnelshift_l(v[c], v[c], 1) & (SLAB_KEY_MASK | 1u)) | ((uint32_t)(c0 + c) << 1);
                        k2 = min(k2, max(k1, key));
                        k1 = min(k1, key);
                    }
                }
            }
            tc_fence_before();
            const uint32_t part_cols = (uint32_t)(part * EPI_COLS) << 1;
            s_part_S[r][part][row] = S;
            s_part_k1[r][part][row] = k1 == 0xFFFFFFFFu ? k1 : k1 + part_cols;   // column bits: local column + first column of this quarter (no carry: < 256)
            s_part_k2[r][part][row] = k2 == 0xFFFFFFFFu ? k2 : k2 + part_cols;
            s_part_zmax[r][part][row] = zmax;
            named_bar_sync(3, SLAB_GEN_THREADS);   // the four column quarters of every row are in
            if (part == 0 && i < p.M) {
                uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;
                S = 0; zmax = 0.0f;
#pragma unroll
                for (int q = 0; q < SLAB_KSPLIT; ++q) {
                    S += s_part_S[r][q][row];
                    zmax += s_part_zmax[r][q][row];   // total sum of squares of the trial's normals in this slab
                    const uint32_t a = s_part_k1[r][q][row], b = s_part_k2[r][q][row];
                    m2 = min(m2, max(m1, a)); m1 = min(m1, a);
                    m2 = min(m2, max(m1, b)); m1 = min(m1, b);
                }
                const float thr = __fmaf_rn(p.delta_rel, sqrtf(zmax) * 1.0001f, p.delta_abs);
                const uint32_t thr_key = __float_as_uint(thr) << 1;
                uint32_t flagged = 0u;
                if ((m1 & SLAB_KEY_MASK) < thr_key) {
                    if ((m2 & SLAB_KEY_MASK) < thr_key || !p.pair_list) {
                        flagged = 0x8000u;           // several near-threshold decisions: the whole trial goes to the fp64 kernel
                    } else {
                        S -= m1 & 1u;                // leave the one uncertain decision out; exact_pairs_kernel adds the fp64 one
                        const uint32_t slot = atomicAdd(p.pair_count, 1u);
                        if (slot < p.pair_cap)
                            p.pair_list[slot] = ((unsigned long long)i << 16) | (unsigned long long)(p.row_base + ((m1 >> 1) & 0xFFu));
                    }
                }
                const uint32_t old = p.first_slab ? 0u : (uint32_t)p.S_acc[i];
                p.S_acc[i] = (uint16_t)(((old & 0x7FFFu) + S) | (old & 0x8000u) | flagged);
            }
        };

        for (uint64_t n = 0; n < my_tiles; ++n) {
            const uint64_t tile = blockIdx.x + (uint64_t)gridDim.x * n;
            const uint64_t i = tile * TILE_M + row;
            const uint64_t path = p.path_offset + i;
            PhiloxPath rng;
            rng.init((uint32_t)path, (uint32_t)(path >> 32), p.key1);
            float zmax = 0.0f;
#pragma unroll 1
            for (int kb = 0; kb < KBN; ++kb) {
                const uint32_t u = (uint32_t)(n * KBN + kb), slot = u & 1;
                if (u >= 2) {
                    mbar_wait(&s_empty[slot], (empty_phase >> slot) & 1u);
                    empty_phase ^= 1u << slot;
                }
                // the slab's transform K-block: 64 KB from the pre-swizzled global image (L2 resident), asynchronous
                {
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.B_slab) + (size_t)kb * SLAB_B_BYTES;
                    const uint32_t dst = smem_u32(smem + slot * SLAB_SLOT_BYTES + SLAB_Z_BYTES);
#pragma unroll
                    for (int q = 0; q < (int)(SLAB_B_BYTES / 16 / SLAB_GEN_THREADS); ++q)
                        cp_async16(dst + 16 * (tid + SLAB_GEN_THREADS * q), src + 16 * (tid + SLAB_GEN_THREADS * q));
                }
                const uint32_t zhi_row = ring_row + slot * SLAB_SLOT_BYTES, zlo_row = zhi_row + TILE_M * 128;
                // this thread's two chunks (16 columns) of the K-block
                const int c0 = 8 * kb + 2 * part;
                if (c0 + 2 <= n_full) {
                    float z[16];
                    rng.normals_multi<4>((uint32_t)(2 * c0), p.key0, p.key1, z);
                    float unused = 0.0f;
#pragma unroll
                    for (int q = 0; q < 16; ++q) zmax = fmaf(z[q], z[q], zmax);   // here "zmax" carries sum z^2
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        float zc[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) zc[q] = z[8 * w + q];
                        store_chunk(zc, 2 * part + w, zhi_row, zlo_row, r7s, unused);
                    }
                } else {
                    for (int c8 = c0; c8 < c0 + 2; ++c8) {   // chunks touching the end of the normal columns
                        float z[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) z[q] = 0.0f;
                        if (8 * c8 < p.ncols) rng.normals((uint32_t)(2 * c8), p.key0, p.key1, z);
                        if (8 * c8 + 4 < p.ncols) rng.normals((uint32_t)(2 * c8 + 1), p.key0, p.key1, z + 4);
                        float zmax_unused = 0.0f;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (8 * c8 + q < p.ncols) zmax = fmaf(z[q], z[q], zmax);
                            if (8 * c8 + q == p.ncols) z[q] = 1.0f;
                            if (8 * c8 + q > p.ncols) z[q] = 0.0f;
                        }
                        store_chunk(z, c8 - 8 * kb, zhi_row, zlo_row, r7s, zmax_unused);
                    }
                }
                cp_async_wait_all();
                fence_async_smem();
                if (kb == 0) tc_fence_before();
                named_bar_arrive(1 + slot, SLAB_THREADS);
            }
            if (n > 0) epilogue(n - 1, prev_i, prev_zmax);
            prev_i = i;
            prev_zmax = zmax;
        }
        if (my_tiles > 0) epilogue(my_tiles - 1, prev_i, prev_zmax);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// reduce the per-trial slab sums: flagged trials go to the fp64 fix-up list, the rest are counted
__global__ void __launch_bounds__(256) slab_count_kernel(uint64_t M, const uint16_t* S_acc, int k, int32_t* S_out,
                                                         uint32_t* flag_list, uint32_t* flag_count, uint32_t flag_cap,
                                                         uint64_t* block_partials) {
    uint32_t cnt = 0, sum_s = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = S_acc[i];
        if (v & 0x8000u) {
            const uint32_t slot = atomicAdd(flag_count, 1u);
            if (slot < flag_cap) flag_list[slot] = (uint32_t)i;
        } else {
            cnt += (v >= (uint32_t)k) ? 1u : 0u;
            sum_s += v;
            if (S_out) S_out[i] = (int32_t)v;
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    sum_s = __reduce_add_sync(0xffffffffu, sum_s);
    __shared__ uint32_t s_c[8], s_s[8];
    if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = cnt; s_s[threadIdx.x >> 5] = sum_s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t c = 0, s = 0;
        for (int w = 0; w < 8; ++w) { c += s_c[w]; s += s_s[w]; }
        block_partials[2 * (uint64_t)blockIdx.x] = c;
        block_partials[2 * (uint64_t)blockIdx.x + 1] = s;
    }
}

// ---- host side ---------------------------------------------------------------------------------------
// Transform image: rows i < N hold A_ij * 2^10 (columns j < N) and (mu_i + x) * 2^10 (column N); padded
// rows hold +2^10 in column N so they are always far on the "no crash" side.  Everything is split into
// fp16 hi + lo.  The 2^10 scale only moves the values (|A| ~ 1e-2, lo parts ~ 1e-6) out of fp16's
// subnormal range; signs and the (equally scaled) guard band are unaffected.
template <int NOUT, int KB>
static void tc_build_image(const double* A, const double* c, int N, const double* port_row, uint8_t* image) {
    const size_t bytes = (size_t)KB * NOUT * 128;
    memset(image, 0, 2 * bytes);
    for (int i = 0; i < NOUT; ++i)
        for (int k = 0; k <= N; ++k) {
            double v;
            if (port_row && i == NOUT - 1) v = port_row[k];   // portfolio: w^T A (k < N) and w.mu (k = N)
            else if (k < N) v = (i < N) ? A[(size_t)i * N + k] : 0.0;
            else v = (i < N) ? c[i] : 1.0;
            v *= TC_SCALE;
            const __half hi = __float2half_rn((float)v);
            const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
            const uint32_t off = tc::operand_offset(NOUT, i, k);
            memcpy(image + off, &hi, 2);
            memcpy(image + bytes + off, &lo, 2);
        }
}

template <int NOUT, int KB, int GROUPS>
static size_t tc_smem_bytes(bool portfolio) {
    return (size_t)GROUPS * 2 * tc::TILE_M * 128 + 2 * (size_t)KB * NOUT * 128 + 1024 + (portfolio ? 8 * (size_t)PORT_BINS : 0);
}

template <int NOUT, int KB, int GROUPS>
static cudaError_t tc_prepare(bool portfolio) {
    // opt in to the large dynamic shared memory and the largest carveout (several CTAs per SM overlap
    // each other's bubbles); both are per-device function attributes, cheap to set repeatedly
    cudaError_t e = cudaFuncSetAttribute(tc_path_kernel<NOUT, KB, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)tc_smem_bytes<NOUT, KB, GROUPS>(portfolio));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc_path_kernel<NOUT, KB, GROUPS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

template <int NOUT, int KB, int GROUPS>
static cudaError_t tc_launch_impl(const TcParams& p, int grid, cudaStream_t s) {
    cudaError_t e = tc_prepare<NOUT, KB, GROUPS>(p.portfolio != 0);
    if (e != cudaSuccess) return e;
    tc_path_kernel<NOUT, KB, GROUPS><<<grid, 128 * GROUPS, tc_smem_bytes<NOUT, KB, GROUPS>(p.portfolio != 0), s>>>(p);
    return cudaGetLastError();
}

// CTAs that fit one SM.  Computed here because cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1
// for this kernel although the hardware co-schedules more (measured: 4 CTAs/SM run 2.2x faster than 1):
// shared memory, TMEM columns (512 per SM), threads and registers bound it.
template <int NOUT, int KB, int GROUPS>
static cudaError_t tc_occupancy_impl(int* bps, bool portfolio) {
    if (tc_smem_bytes<NOUT, KB, GROUPS>(portfolio) > 227 * 1024) { *bps = 0; return cudaSuccess; }
    cudaError_t e = tc_prepare<NOUT, KB, GROUPS>(portfolio);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa{};
    e = cudaFuncGetAttributes(&fa, tc_path_kernel<NOUT, KB, GROUPS>);
    if (e != cudaSuccess) return e;
    int dev = 0, smem_sm = 0, regs_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
    constexpr int threads = 128 * GROUPS;
    constexpr uint32_t group_cols = (NOUT + 31) & ~31, need = group_cols * GROUPS;
    constexpr int tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    const size_t per_cta = tc_smem_bytes<NOUT, KB, GROUPS>(portfolio) + fa.sharedSizeBytes + 1024;  // + driver reservation
    int n = (int)(smem_sm / per_cta);
    n = std::min(n, 512 / tmem_cols);
    n = std::min(n, 2048 / threads);
    n = std::min(n, regs_sm / (((fa.numRegs + 7) & ~7) * ((threads + 31) & ~31)));
    *bps = std::max(n, 0);
    if (getenv("MCR_DEBUG"))
        fprintf(stderr, "[mcr] tc_path_kernel<%d,%d,%d>: %d CTAs/SM (dyn smem %zu, static %zu, regs %d, tmem cols %d)\n",
                NOUT, KB, GROUPS, *bps, tc_smem_bytes<NOUT, KB, GROUPS>(portfolio), fa.sharedSizeBytes, fa.numRegs, tmem_cols);
    return cudaSuccess;
}

struct TcVariant {
    int nout, kb, groups;
    size_t image_bytes;
    void (*build)(const double*, const double*, int, const double*, uint8_t*);
    cudaError_t (*launch)(const TcParams&, int, cudaStream_t);
    cudaError_t (*occupancy)(int*, bool);
};

template <int NOUT, int KB, int GROUPS>
static TcVariant tc_variant() {
    return TcVariant{NOUT, KB, GROUPS, 2 * (size_t)KB * NOUT * 128, &tc_build_image<NOUT, KB>, &tc_launch_impl<NOUT, KB, GROUPS>,
                     &tc_occupancy_impl<NOUT, KB, GROUPS>};
}

// warpgroups per CTA: as many as the tuning default, TMEM (512 columns) and shared memory (227 KB) allow
constexpr int tc_groups(int nout, int kb) {
    int g = MCR_TC_GROUPS;
    const int cols = (nout + 31) & ~31;
    while (g > 1 && (g * cols > 512 || g * 32768 + 2 * kb * nout * 128 + 4096 > 227 * 1024)) --g;
    return g;
}

// smallest instantiation with nout >= N and 64 * kb >= N + 1
static const TcVariant* tc_find(int N) {
    static const TcVariant table[] = {
        tc_variant<16, 1, tc_groups(16, 1)>(), tc_variant<32, 1, tc_groups(32, 1)>(), tc_variant<48, 1, tc_groups(48, 1)>(),
        tc_variant<64, 1, tc_groups(64, 1)>(), tc_variant<64, 2, tc_groups(64, 2)>(), tc_variant<80, 2, tc_groups(80, 2)>(),
        tc_variant<96, 2, tc_groups(96, 2)>(), tc_variant<112, 2, tc_groups(112, 2)>(), tc_variant<128, 2, tc_groups(128, 2)>(),
        tc_variant<128, 3, tc_groups(128, 3)>(),
    };
    for (const TcVariant& v : table)
        if (v.nout >= N && 64 * v.kb >= N + 1) return &v;
    return nullptr;
}

int tc_max_n() { return 128; }

bool tc_supported(int rows) { return tc_find(rows) != nullptr; }

size_t tc_image_bytes(int rows) { const TcVariant* v = tc_find(rows); return v ? v->image_bytes : 0; }

int tc_padded_n(int rows) { const TcVariant* v = tc_find(rows); return v ? v->nout : 0; }

void tc_build(int N, int rows, const double* A, const double* c, const double* port_row, uint8_t* image) { tc_find(rows)->build(A, c, N, port_row, image); }

cudaError_t tc_occupancy(int rows, bool portfolio, int* bps) { return tc_find(rows)->occupancy(bps, portfolio); }

cudaError_t tc_launch(int N, int rows, const TcLaunch& l, int grid, cudaStream_t s) {
    TcParams p{};
    p.M = l.M; p.path_offset = l.path_offset; p.key0 = l.key0; p.key1 = l.key1; p.k = l.k; p.N = N;
    p.delta_rel = (float)(l.delta_rel * TC_SCALE); p.delta_abs = (float)(l.delta_abs * TC_SCALE);  // accumulators are scaled
    p.B_image = reinterpret_cast<const uint4*>(l.d_image); p.S_out = l.S_out;
    p.flag_list = l.flag_list; p.flag_count = l.flag_count; p.flag_cap = l.flag_cap; p.block_partials = l.block_partials;
    p.pair_list = l.pair_list; p.pair_count = l.pair_count; p.pair_cap = l.pair_cap;
    p.portfolio = l.portfolio; p.port_thresh = (float)(-l.port_x * TC_SCALE);
    p.hist_lo = (float)l.hist_lo; p.hist_inv_w = (float)(1.0 / l.hist_width);
    p.port_acc = l.port_acc; p.port_hist = l.port_hist; p.port_hist_sum = l.port_hist_sum;
    for (int q = 0; q < 8; ++q) {
        const int col = 8 * (N / 8) + q;
        p.tail_mul[q] = col < N ? 1.0f : 0.0f;
        p.tail_add[q] = col == N ? 1.0f : 0.0f;
    }
    return tc_find(rows)->launch(p, grid, s);
}

// ---- slab path host side ---------------------------------------------------------------------------------
int slab_rows() { return SLAB_ROWS; }
int slab_count(int N) { return (N + SLAB_ROWS - 1) / SLAB_ROWS; }
int slab_ncols(int N, int t, bool lower) { return lower ? std::min(N, SLAB_ROWS * (t + 1)) : N; }
int slab_kblocks(int ncols) { return (ncols + 1 + tc::KBLOCK - 1) / tc::KBLOCK; }

size_t slab_image_bytes(int N, bool lower) {
    size_t b = 0;
    for (int t = 0; t < slab_count(N); ++t) b += (size_t)slab_kblocks(slab_ncols(N, t, lower)) * SLAB_B_BYTES;
    return b;
}

// image of every slab, back to back; offsets[t] = byte offset of slab t (offsets has slab_count(N) + 1 entries)
void slab_build(int N, bool lower, const double* A, const double* c, uint8_t* image, size_t* offsets) {
    size_t off = 0;
    for (int t = 0; t < slab_count(N); ++t) {
        offsets[t] = off;
        const int ncols = slab_ncols(N, t, lower), kbn = slab_kblocks(ncols);
        uint8_t* img = image + off;
        memset(img, 0, (size_t)kbn * SLAB_B_BYTES);
        for (int r = 0; r < SLAB_ROWS; ++r) {
            const int i = SLAB_ROWS * t + r;
            for (int k = 0; k <= ncols; ++k) {
                double v;
                if (k < ncols) v = (i < N) ? A[(size_t)i * N + k] : 0.0;
                else v = (i < N) ? c[i] : 1.0;
                v *= TC_SCALE;
                const __half hi = __float2half_rn((float)v);
                const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
                // K-block kb = k / 64: [hi: 128 rows x 128 B][lo: ...], same swizzle as the device-side Z rows
                const int kb = k / tc::KBLOCK, kk = k % tc::KBLOCK;
                const size_t o = (size_t)kb * SLAB_B_BYTES + (size_t)(r / 8) * 1024 + (size_t)(r % 8) * 128 +
                                 (size_t)(((kk / 8) ^ (r % 8)) * 16) + (size_t)(kk % 8) * 2;
                memcpy(img + o, &hi, 2);
                memcpy(img + o + SLAB_ROWS * 128, &lo, 2);
            }
        }
        off += (size_t)kbn * SLAB_B_BYTES;
    }
    offsets[slab_count(N)] = off;
}

cudaError_t slab_launch(const SlabLaunch& l, int num_sms, cudaStream_t s) {
    const size_t smem = 2 * (size_t)SLAB_SLOT_BYTES + 1024;
    cudaError_t e = cudaFuncSetAttribute(tc_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    TcSlabParams p{};
    p.M = l.M; p.path_offset = l.path_offset; p.key0 = l.key0; p.key1 = l.key1;
    p.ncols = l.ncols; p.kb_count = slab_kblocks(l.ncols);
    p.delta_rel = (float)(l.delta_rel * TC_SCALE); p.delta_abs = (float)(l.delta_abs * TC_SCALE);
    p.thr_max = p.delta_rel * 7.0f * sqrtf((float)l.ncols) * 1.001f + p.delta_abs;
    p.B_slab = reinterpret_cast<const uint4*>(l.d_slab_image); p.S_acc = l.S_acc; p.first_slab = l.first_slab;
    p.row_base = l.row_base; p.pair_list = l.pair_list; p.pair_count = l.pair_count; p.pair_cap = l.pair_cap;
    const uint64_t tiles = (l.M + tc::TILE_M - 1) / tc::TILE_M;
    const int grid = (int)std::min<uint64_t>((uint64_t)num_sms, tiles ? tiles : 1);
    tc_slab_kernel<<<grid, SLAB_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t slab_count_launch(uint64_t M, const uint16_t* S_acc, int k, int32_t* S_out, uint32_t* flag_list,
                              uint32_t* flag_count, uint32_t flag_cap, uint64_t* block_partials, int grid, cudaStream_t s) {
    slab_count_kernel<<<grid, 256, 0, s>>>(M, S_acc, k, S_out, flag_list, flag_count, flag_cap, block_partials);
    return cudaGetLastError();
}

}  // namespace mcr
