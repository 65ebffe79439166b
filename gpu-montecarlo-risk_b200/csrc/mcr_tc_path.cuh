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
//
// This header holds the kernel template and its host-side launch helpers; mcr_tc_inst.cu instantiates one
// (NOUT, KB) variant per translation unit (the Makefile compiles it once per MCR_TC_INST value, in parallel),
// mcr_tc.cu looks the variants up.
#pragma once
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
    PhiloxKeys keys;              // the round keys of (key0, key1): constant-bank operands of the generator
    int32_t k, N;
    float delta_rel, delta_abs;   // guard band = delta_rel * max_j |z_j| + delta_abs (0, 0 disables the recheck)
    const uint4* B_image;   // global: [Bhi image | Blo image], each KB * NOUT * 128 bytes
    const double* Z;        // SUPPLIED instantiation: the caller's normals, M x N fp64 row-major (chunk-relative)
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
    float port_thresh;            // scaled: tail event is  D < port_thresh  (= -x_p * scale of the portfolio row)
    float port_inv_scale;         // 1 / scale of the portfolio row
    float hist_lo, hist_inv_w;    // unscaled histogram origin and 1 / bin width
    double hist_lo_d, hist_w_d;   // the same origin and the bin width in fp64 (per-bin sums are rebuilt from the quantised positions)
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

// MCR_TC_FOLD = 1: the generator emits z' = z / sqrt(2 ln 2) (box_muller_raw) and the transform image carries the factor
// instead; the guard band then uses the largest Box-Muller radius of the trial as its bound on |z'|.  Saves 3 of the
// ~82 instructions per Philox block.  The normals of the fp64 fix-up / dump kernels are unchanged: fl32(s * z') differs
// from them by a few ulp, far inside the band (2^-19 relative, doubled).
#ifndef MCR_TC_FOLD
#define MCR_TC_FOLD 1
#endif
constexpr bool TC_FOLD = MCR_TC_FOLD != 0;
constexpr double TC_Z_FACTOR = TC_FOLD ? BOX_MULLER_S : 1.0;   // what the transform columns 0..N-1 are multiplied by

// NB Philox blocks b.. of the path -> 4 NB normals (only LAST_NORMALS of the last block), and the running bound on |z|.
// SUPPLIED: the 4 NB normals of columns 4 b .. come from the caller's fp64 row instead (zrow = nullptr for a trial past the end):
// parity mode 1 "fed identical pre-generated normals through the interface" for this kernel.
template <int NB, int LAST_NORMALS = 4, bool SUPPLIED = false>
__device__ __forceinline__ void gen_blocks(const PhiloxPath& rng, uint32_t b, const PhiloxKeys& K, float* z, float& zbound,
                                           const double* zrow = nullptr, int ncols = 0) {
    if (SUPPLIED) {
#pragma unroll
        for (int q = 0; q < 4 * NB; ++q) {
            const int col = 4 * (int)b + q;
            float v = (zrow != nullptr && col < ncols) ? __double2float_rn(zrow[col]) : 0.0f;
            if (TC_FOLD) v = __fmul_rn(v, (float)(1.0 / BOX_MULLER_S));   // the image carries the factor back (one more fp32 rounding, inside the band)
            z[q] = v;
            zbound = fmaxf(zbound, fabsf(v));
        }
        return;
    }
    if (TC_FOLD) {
        rng.normals_multi_raw<NB, LAST_NORMALS>(b, K, z, zbound);
    } else {   // zbound is tracked by store_chunk<true>
        if (NB == 1 && LAST_NORMALS == 2) rng.normals_pair(b, K, z);
        else if (NB == 1) rng.normals(b, K, z);
        else rng.normals_multi<NB>(b, K, z);
    }
}


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
template <int NOUT, int KB, int GROUPS, bool SUPPLIED = false>
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
    // only when p.portfolio (host sizes smem): one 64-bit word per bin = count << 44 | sum of the trials' positions in the histogram
    // range quantised to 2^-24 of it (port_hist_word): integer adds, exact and independent of the order the trials arrive in
    unsigned long long* s_hist = reinterpret_cast<unsigned long long*>(s_blo + B_BYTES);
    __shared__ uint64_t s_mma[GROUPS];
    __shared__ double s_port[4 * GROUPS][4];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_cnt[4 * GROUPS], s_sum[4 * GROUPS];

    // warp index through a shuffle: the compiler then keeps everything derived from it (group, slot and TMEM
    // addresses, the UMMA descriptors) in uniform registers
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

    // ---- one-time setup ----
    {
        const uint4* src = p.B_image;
        uint4* dst = reinterpret_cast<uint4*>(s_bhi);
        for (uint32_t t = tid; t < 2 * B_BYTES / 16; t += NTHREADS) dst[t] = src[t];
    }
    if (p.portfolio)
        for (uint32_t t = tid; t < PORT_BINS; t += NTHREADS) s_hist[t] = 0ull;
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
        const uint64_t zhi_desc = make_desc(slot_base), zlo_desc = make_desc(slot_base + TILE_M * 128);
        const uint64_t bhi_desc = make_desc(smem_u32(s_bhi)), blo_desc = make_desc(smem_u32(s_blo));
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
                            const float P = Ps * p.port_inv_scale;
                            const bool tail = Ps < p.port_thresh;
                            p_hits += tail ? 1u : 0u;
                            p_tail += tail ? -P : 0.0f;
                            p_sum += P;
                            p_sum2 = fmaf(P, P, p_sum2);
                            int bin;
                            const unsigned long long word = port_hist_word(P, p.hist_lo, p.hist_inv_w, bin);
                            atomicAdd(&s_hist[bin], word);
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
            const double* zrow = (SUPPLIED && i < p.M) ? p.Z + (size_t)i * (size_t)p.N : nullptr;
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
                    gen_blocks<2 * TC_UNROLL, 4, SUPPLIED>(rng, (uint32_t)(2 * c8), p.keys, z, zmax, zrow, p.N);
                    acquire();
#pragma unroll
                    for (int w = 0; w < TC_UNROLL; ++w) {
                        float zc[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) zc[q] = z[8 * w + q];
                        store_chunk<!TC_FOLD>(zc, c8 + w - c_begin, zhi_row, zlo_row, r7s, zmax);
                    }
                }
                for (; c8 < min(n_full, c_begin + 8); ++c8) {
                    float z[8];
                    gen_blocks<2, 4, SUPPLIED>(rng, (uint32_t)(2 * c8), p.keys, z, zmax, zrow, p.N);
                    acquire();
                    store_chunk<!TC_FOLD>(z, c8 - c_begin, zhi_row, zlo_row, r7s, zmax);
                }
                acquire();
                if (n_full < c_begin + 8) {
                    // last K-block of the tile: the chunk holding the remaining rem = N % 8 normals and the constant 1 in
                    // column N, then all-zero chunks.  Only the Philox blocks / Box-Muller pairs that are needed are drawn.
                    const int rem = p.N - 8 * n_full;
                    float z[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) z[q] = 0.0f;
                    if (rem > 6) gen_blocks<2, 4, SUPPLIED>(rng, (uint32_t)(2 * n_full), p.keys, z, zmax, zrow, p.N);
                    else if (rem > 4) gen_blocks<2, 2, SUPPLIED>(rng, (uint32_t)(2 * n_full), p.keys, z, zmax, zrow, p.N);
                    else if (rem > 2) gen_blocks<1, 4, SUPPLIED>(rng, (uint32_t)(2 * n_full), p.keys, z, zmax, zrow, p.N);
                    else if (rem > 0) gen_blocks<1, 2, SUPPLIED>(rng, (uint32_t)(2 * n_full), p.keys, z, zmax, zrow, p.N);
#pragma unroll
                    for (int q = 0; q < 8; ++q)   // mul/add are (1, 0) for the normals, (0, 1) for the constant 1 in column N, (0, 0) beyond:
                        z[q] = __fmaf_rn(z[q], p.tail_mul[q], p.tail_add[q]);   // exact; the 1 only widens zmax (never below 1)
                    store_chunk<!TC_FOLD>(z, n_full - c_begin, zhi_row, zlo_row, r7s, zmax);
                    // the zero chunks of a single-K-block tile sit at the same place in every tile of the group
                    if (KB > 1 || n == 0) {
#pragma unroll 1
                        for (c8 = n_full + 1; c8 < c_begin + 8; ++c8) store_zero_chunk(c8 - c_begin, zhi_row, zlo_row, r7s);
                    }
                }
                // phase u - 1 is complete here; when it closed tile n - 1, that tile's accumulator is final
                if (kb == 0 && n > 0) epilogue(prev_i, prev_zmax);
                fence_async_smem();     // generic-proxy smem writes -> visible to the tensor core
                tc_fence_before();      // our TMEM reads are ordered before the hand-over
                if (wg == (int)(u & 3u)) {
                    named_bar_sync(1 + g, 128);   // the K-block is complete and everyone is done with the accumulator
                    if (elect_one()) {   // a single thread as far as the compiler is concerned: no ELECT retry loop around each UTCHMMA
                        tc_fence_after();
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
                            // descriptors of the 16-column K-steps: start address field + 2 (32 bytes >> 4) per step
                            const uint64_t za = pass == 2 ? zlo_desc : zhi_desc;
                            const uint64_t ba = (pass == 1 ? blo_desc : bhi_desc) + (uint64_t)(kb * (NOUT * 128 / 16));
#pragma unroll
                            for (int ks = 0; ks < (MCR_TC_ABLATE == 3 ? 0 : KBLOCK / 16); ++ks)
                                umma_f16(d_tmem, za + 2 * ks, ba + 2 * ks, IDESC, (kb | pass | ks) != 0 ? 1u : 0u);
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
            port_hist_flush(s_hist[t], t, p.hist_lo_d, p.hist_w_d, p.port_hist, p.port_hist_sum);
        }
    }
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---- host side ---------------------------------------------------------------------------------------
// Transform image: row i < N holds s_i * A_ij (columns j < N) and s_i * (mu_i + x) (column N), s_i = the row's power-of-two
// scale (tc_row_scale; returned in row_scale[NOUT]); padded rows hold a large positive constant in column N so they are
// always far on the "no crash" side.  Everything is split into fp16 hi + lo.  The scale only moves the values into the
// middle of fp16's normal range; signs and the (equally scaled) guard band are unaffected.
template <int NOUT, int KB>
static void tc_build_image(const double* A, const double* c, int N, const double* port_row, uint8_t* image, double* row_scale) {
    const size_t bytes = (size_t)KB * NOUT * 128;
    memset(image, 0, 2 * bytes);
    for (int i = 0; i < NOUT; ++i) {
        const bool port = port_row && i == NOUT - 1;
        double mx = 0.0;
        for (int k = 0; k <= N; ++k) {
            const double v = port ? port_row[k] : (i < N ? (k < N ? A[(size_t)i * N + k] : c[i]) : 0.0);
            mx = std::max(mx, std::fabs(v) * (k < N ? TC_Z_FACTOR : 1.0));
        }
        row_scale[i] = (port || i < N) ? tc_row_scale(mx) : 1.0;
    }
    for (int i = 0; i < NOUT; ++i)
        for (int k = 0; k <= N; ++k) {
            double v;
            if (port_row && i == NOUT - 1) v = port_row[k];   // portfolio: w^T A (k < N) and w.mu (k = N)
            else if (k < N) v = (i < N) ? A[(size_t)i * N + k] : 0.0;
            else v = (i < N) ? c[i] : TC_PAD_ROW;
            v *= row_scale[i] * (k < N ? TC_Z_FACTOR : 1.0);
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

template <int NOUT, int KB, int GROUPS, bool SUPPLIED = false>
static cudaError_t tc_prepare(bool portfolio) {
    // opt in to the large dynamic shared memory and the largest carveout (several CTAs per SM overlap
    // each other's bubbles); both are per-device function attributes, cheap to set repeatedly
    cudaError_t e = cudaFuncSetAttribute(tc_path_kernel<NOUT, KB, GROUPS, SUPPLIED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)tc_smem_bytes<NOUT, KB, GROUPS>(portfolio));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc_path_kernel<NOUT, KB, GROUPS, SUPPLIED>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

template <int NOUT, int KB, int GROUPS, bool SUPPLIED = false>
static cudaError_t tc_launch_impl(const TcParams& p, int grid, cudaStream_t s) {
    cudaError_t e = tc_prepare<NOUT, KB, GROUPS, SUPPLIED>(p.portfolio != 0);
    if (e != cudaSuccess) return e;
    tc_path_kernel<NOUT, KB, GROUPS, SUPPLIED><<<grid, 128 * GROUPS, tc_smem_bytes<NOUT, KB, GROUPS>(p.portfolio != 0), s>>>(p);
    return cudaGetLastError();
}

// CTAs that fit one SM.  Computed here because cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1
// for this kernel although the hardware co-schedules more (measured: 4 CTAs/SM run 2.2x faster than 1):
// shared memory, TMEM columns (512 per SM), threads and registers bound it.
template <int NOUT, int KB, int GROUPS, bool SUPPLIED = false>
static cudaError_t tc_occupancy_impl(int* bps, bool portfolio) {
    if (tc_smem_bytes<NOUT, KB, GROUPS>(portfolio) > 227 * 1024) { *bps = 0; return cudaSuccess; }
    cudaError_t e = tc_prepare<NOUT, KB, GROUPS, SUPPLIED>(portfolio);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa{};
    e = cudaFuncGetAttributes(&fa, tc_path_kernel<NOUT, KB, GROUPS, SUPPLIED>);
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
    void (*build)(const double*, const double*, int, const double*, uint8_t*, double*);
    cudaError_t (*launch)(const TcParams&, int, cudaStream_t);
    cudaError_t (*occupancy)(int*, bool);
};

template <int NOUT, int KB, int GROUPS, bool SUPPLIED = false>
static TcVariant tc_variant() {
    return TcVariant{NOUT, KB, GROUPS, 2 * (size_t)KB * NOUT * 128, &tc_build_image<NOUT, KB>, &tc_launch_impl<NOUT, KB, GROUPS, SUPPLIED>,
                     &tc_occupancy_impl<NOUT, KB, GROUPS, SUPPLIED>};
}

// warpgroups per CTA: as many as the tuning default, TMEM (512 columns) and shared memory (227 KB) allow
constexpr int tc_groups(int nout, int kb) {
    int g = MCR_TC_GROUPS;
    const int cols = (nout + 31) & ~31;
    while (g > 1 && (g * cols > 512 || g * 32768 + 2 * kb * nout * 128 + 4096 > 227 * 1024)) --g;
    return g;
}

// the (NOUT, KB) instantiations, smallest first; variant i is defined by mcr_tc_inst.cu compiled with -DMCR_TC_INST=i
constexpr int TC_VARIANTS = 10;
constexpr int TC_VARIANT_NOUT[TC_VARIANTS] = {16, 32, 48, 64, 64, 80, 96, 112, 128, 128};
constexpr int TC_VARIANT_KB[TC_VARIANTS] = {1, 1, 1, 1, 2, 2, 2, 2, 2, 3};
template <int I> TcVariant tc_variant_inst();
// the one supplied-normals instantiation (translation unit MCR_TC_INST = TC_VARIANTS): the widest shape, it covers every N <= 128
constexpr int TC_SUPPLIED_INST = TC_VARIANTS, TC_SUPPLIED_NOUT = 128, TC_SUPPLIED_KB = 3;

}  // namespace mcr
