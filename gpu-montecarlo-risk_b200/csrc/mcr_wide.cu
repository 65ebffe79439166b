// Single-pass tensor-core kernel for N > 128 ("wide" portfolios: the Stoxx600 shape, up to N = 1024).
//
// The row-slab kernel (mcr_tc.cu) keeps one slab of accumulators in TMEM and therefore re-derives the
// normals once per slab.  Here the GEMM is turned around so that EVERY accumulator of a tile of trials is
// resident in TMEM at once and each normal is generated exactly once:
//
//     D_t[128 rows x NT trials] += B_{t,kb}[128 rows x 64] * Z_kb[NT trials x 64]^T      (tcgen05.mma, M = 128, N = NT)
//
//   * t = 128-row slab of the transform (at most 8), kb = 64-column K-block, NT = 128 / 96 / 64 trials per tile
//     so that nslab * NT <= 512 TMEM columns;
//   * the K loop is the outer one: the generator warps produce Z_kb (Philox -> Box-Muller -> fp16 hi + lo,
//     swizzled, double-buffered in shared memory), a loader thread streams the transform blocks B_{t,kb}
//     (32 KB each, pre-split and pre-swizzled on the host, in consumption order, L2-resident) through a ring with
//     cp.async.bulk, and the MMA warp issues 12 MMAs (fp16x3, four K-steps) per block.  With a lower-triangular
//     transform, K-block kb is only multiplied into the slabs t >= kb / 2;
//   * the constant term does not ride in an extra K column here: the accumulators of the next tile are
//     initialised with (mu_i + x) * 2^10 through tcgen05.st while the tile's first K-blocks are generated;
//   * epilogue: TMEM lane = asset, column = trial.  Per 32 x 32 piece every lane packs the sign bits of its row, a
//     shuffle bit-transpose gives lane c the column of trial c (popc = its hits), and the smallest |accumulator|
//     per lane tells whether the piece can hold a near-threshold decision at all; only then is it examined
//     decision by decision.  Trials with exactly one
//     near-threshold decision go to exact_pairs_kernel (that row alone, fp64), trials with several to the fp64 kernel.
#include "../../include/mcr_b200.h"
#include "mcr_internal.h"
#include "mcr_rng.cuh"
#include "mcr_tc_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace mcr {

// tuning only (tools/devbuild.sh): 1 = the loader copies each ring slot once and then only signals it (no transform stream),
// 2 = the generators skip Philox / Box-Muller (stale Z), 3 = the MMA warp commits without issuing MMAs.  Results are garbage.
#ifndef MCR_WIDE_ABLATE
#define MCR_WIDE_ABLATE 0
#endif
// Generator / epilogue threads per CTA (+ 64: one MMA-issue warp, one transform-loader warp).  PARTS = the largest power of two
// <= threads / NT of them share a trial's K-block, so with 512 threads a 96-trial tile (513 <= N <= 640) keeps only 384 busy
// generating; MCR_WIDE_GEN_THREADS_96 = 768 gives it 8 parts of 8 columns (build option, to be measured in round 2; 72 registers
// per thread then).
#ifndef MCR_WIDE_GEN_THREADS_96
#define MCR_WIDE_GEN_THREADS_96 512
#endif
__host__ __device__ constexpr int wide_gen_threads(int nt) { return nt == 96 ? MCR_WIDE_GEN_THREADS_96 : 512; }
__host__ __device__ constexpr int wide_parts(int nt) { return wide_gen_threads(nt) / nt >= 8 ? 8 : 4; }
// Shared memory is split between Z slots (K-blocks of normals the generators may be ahead of the tensor core) and the ring of
// transform blocks in flight from L2.  Z slot = NT * 256 B, transform block = 32 KB; ~221 KB usable next to the static part.
// Measured (profiles/r01_wide_zslots_early_epilogue.log, r01_mma_issue_fix.log): more than two Z slots buy nothing, so the
// ring gets the memory.
#ifndef MCR_WIDE_ZSLOTS_128
#define MCR_WIDE_ZSLOTS_128 2
#endif
#ifndef MCR_WIDE_RING_128
#define MCR_WIDE_RING_128 4
#endif
#ifndef MCR_WIDE_ZSLOTS_96
#define MCR_WIDE_ZSLOTS_96 2
#endif
#ifndef MCR_WIDE_RING_96
#define MCR_WIDE_RING_96 5
#endif
#ifndef MCR_WIDE_ZSLOTS_64
#define MCR_WIDE_ZSLOTS_64 2
#endif
#ifndef MCR_WIDE_RING_64
#define MCR_WIDE_RING_64 5
#endif
// Lower-triangular transform: with MCR_WIDE_ELAG = E > 0 slab t (last MMAs in K-block 2t + 1) is thresholded and re-initialised
// inside the K loop, before the generators start K-block 2t + 1 + E; 0 = every slab at the tile boundary.  Correct either way
// (tests run with both), but measured slower with E = 2 / 3 (N = 320: 14.6 / 13.6 ms against 12.7 per 2e7 trials, N = 600:
// 32.4 / 31.8 against 30.8; before the MMA-issue fix the order was the reverse by 2 %): the ncu capture shows the generators
// waiting for the slab's MMAs in the middle of the tile (29 % of all warp samples), time they do not lose at the tile boundary,
// where the tensor core already has two K-blocks of the next tile queued.
#ifndef MCR_WIDE_ELAG
#define MCR_WIDE_ELAG 0
#endif
__host__ __device__ constexpr int wide_zslots(int nt) { return nt >= 128 ? MCR_WIDE_ZSLOTS_128 : nt >= 96 ? MCR_WIDE_ZSLOTS_96 : MCR_WIDE_ZSLOTS_64; }
__host__ __device__ constexpr int wide_ring(int nt) { return nt >= 128 ? MCR_WIDE_RING_128 : nt >= 96 ? MCR_WIDE_RING_96 : MCR_WIDE_RING_64; }
constexpr int WIDE_BAR_ACC = 8, WIDE_BAR_GEN = 9;         // named barriers: 1 .. ZS = Z slot full, accumulators ready, generator sync
constexpr uint32_t WIDE_BLOCK_BYTES = 2 * 128 * 128;       // 128 rows x 64 columns, fp16 hi + lo

struct WideParams {
    uint64_t M, path_offset;
    uint32_t key0, key1;
    int32_t N, k;
    int32_t kbn;              // K-blocks: ceil(N / 64)
    int32_t nslab;            // 128-row slabs: ceil(N / 128)
    int32_t lower;            // K-block kb only feeds the slabs t >= kb / 2
    const uint8_t* image;     // transform blocks in consumption order (kb outer, t inner)
    const float* c_scaled;    // [nslab * 128]: (mu_i + x) * 2^10, padded rows +2^10
    float delta_rel[WIDE_MAX_SLABS], delta_abs[WIDE_MAX_SLABS];   // scaled guard band of slab t: delta_rel * ||z||_2 + delta_abs
    int32_t* S_out;
    uint32_t* flag_list;
    uint32_t* flag_count;
    uint32_t flag_cap;
    unsigned long long* pair_list;   // trial << 21 | other hits << 10 | row, or nullptr
    uint32_t* pair_count;
    uint32_t pair_cap;
    uint64_t* block_partials;
};

namespace tc {
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (TMA, no tensor map: the block is contiguous and already in its shared-memory layout)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns <- the same register value in every column
__device__ __forceinline__ void tmem_st32_bcast(uint32_t taddr, uint32_t r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(r)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
}  // namespace tc

// one stage of the 32 x 32 bit-matrix transpose held one row per lane: swap the off-diagonal J x J blocks with lane ^ J
template <int J, uint32_t MASK>
__device__ __forceinline__ uint32_t transpose_step(uint32_t x, int lane) {
    const uint32_t t = __shfl_xor_sync(0xffffffffu, x, J);
    return (lane & J) ? ((x & MASK) | ((t & MASK) >> J)) : ((x & ~MASK) | ((t & ~MASK) << J));
}

template <int NT>
__global__ void __launch_bounds__(wide_gen_threads(NT) + 64, 1) tc_wide_kernel(const WideParams p) {
    using namespace tc;
    constexpr int WIDE_GEN_THREADS = wide_gen_threads(NT);
    constexpr int PARTS = wide_parts(NT);              // generator threads per trial
    constexpr int QWARPS = WIDE_GEN_THREADS / 128;     // warps per TMEM lane quarter
    static_assert(PARTS * NT <= WIDE_GEN_THREADS && WIDE_GEN_THREADS % 128 == 0, "generator thread layout");
    constexpr int PCOLS = KBLOCK / PARTS;              // columns per generator thread and K-block
    constexpr uint32_t ZSLOT = 2 * NT * 128;           // one K-block of Z: hi then lo
    constexpr uint32_t IDESC = make_idesc(NT);
    constexpr int G32 = NT / 32;                       // 32-trial column groups
    constexpr int ALL = WIDE_GEN_THREADS + 32;         // generators + MMA warp (named barriers 1..3)
    constexpr int MMA_WARP = WIDE_GEN_THREADS / 32, LOAD_WARP = MMA_WARP + 1;
    constexpr int WIDE_RING = wide_ring(NT), ZS = wide_zslots(NT);
    static_assert(ZS >= 2 && ZS <= 7, "Z slots use named barriers 1 .. ZS");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* s_z = smem;                   // [ZS][hi | lo][NT x 128 B]
    uint8_t* s_ring = smem + ZS * ZSLOT;   // [WIDE_RING][hi | lo][128 x 128 B]
    __shared__ uint64_t s_zempty[ZS], s_bfull[WIDE_RING], s_bempty[WIDE_RING], s_done, s_slabdone[WIDE_MAX_SLABS];
    __shared__ uint32_t s_tmem;
    __shared__ float s_zsq[8][128];
    __shared__ uint32_t s_S[128], s_namb[128], s_amb[128];
    __shared__ uint32_t s_cnt[4], s_sum[4];

    // the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role's loop state (ring slot,
    // shared-memory addresses, UMMA descriptors) in uniform registers; as a plain tid >> 5 the issuing lane spent ~16
    // instructions per MMA (ELECT + R2UR.BROADCAST x5 + a retry branch) moving them there and was the kernel's critical path
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        for (int b = 0; b < ZS; ++b) mbar_init(&s_zempty[b], 1);
        for (int b = 0; b < WIDE_RING; ++b) { mbar_init(&s_bfull[b], 1); mbar_init(&s_bempty[b], 1); }
        for (int b = 0; b < WIDE_MAX_SLABS; ++b) mbar_init(&s_slabdone[b], 1);
        mbar_init(&s_done, 1);
    }
    if (tid < 128) { s_S[tid] = 0u; s_namb[tid] = 0u; s_amb[tid] = 0u; }
    if (warp == MMA_WARP) tmem_alloc(&s_tmem, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    const uint64_t n_tiles = (p.M + NT - 1) / NT;
    const uint64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int KBN = p.kbn, NSLAB = p.nslab;
    uint32_t cnt = 0, sum_s = 0;
    // With a lower-triangular transform slab t receives its last MMAs in K-block 2t + 1 (the MMA warp then commits
    // s_slabdone[t]).  Slabs 0 .. n_early - 1 are thresholded and re-initialised by the generator warps inside the K loop,
    // before they start K-block 2t + 1 + ELAG, while the tensor core works on the other slabs; the remaining slab(s) are left
    // for the tile boundary, after the generators have queued K-blocks 0 .. kb_boundary of the next tile.  The MMA warp starts
    // the next tile on the early slabs (their re-initialisation precedes, in the program order of every generator thread,
    // the hand-over of the next tile's first Z K-block) and waits for the boundary slabs only when it reaches them.
    constexpr int ELAG = MCR_WIDE_ELAG;
    const int n_early = (ELAG > 0 && p.lower && KBN - 2 - ELAG >= 0) ? min((KBN - 2 - ELAG) / 2 + 1, NSLAB - 1) : 0;
    const int kb_boundary = min(min(ZS - 1, ELAG > 0 ? ELAG : ZS - 1), KBN - 1);

    if (warp == MMA_WARP) {
        // ---- MMA issue: K-blocks in order; for each, the slabs that need it, as their transform blocks arrive ----
        uint32_t ukb = 0, blk = 0;
        for (uint64_t n = 0; n < my_tiles; ++n) {
            for (int kb = 0; kb < KBN; ++kb, ++ukb) {
                const uint32_t slot = ukb % ZS;
                named_bar_sync(1 + slot, ALL);   // Z K-block complete and visible to the async proxy
                // descriptors once per K-block / per transform block; the 16-column K-steps add 2 (32 bytes >> 4) to the address field
                const uint64_t zhi_d = make_desc(smem_u32(s_z + slot * ZSLOT)), zlo_d = zhi_d + (uint64_t)(NT * 128 / 16);
                for (int t = p.lower ? (kb >> 1) : 0; t < NSLAB; ++t, ++blk) {
                    // slabs n_early .. : initialised (and the previous tile's epilogue is done with them)
                    if (kb == 0 && t == n_early) named_bar_sync(WIDE_BAR_ACC, ALL);
                    const uint32_t rs = blk % WIDE_RING;
                    mbar_wait(&s_bfull[rs], (blk / WIDE_RING) & 1u);
                    if (elect_one()) {
                        tc_fence_after();
                        const uint64_t bhi_d = make_desc(smem_u32(s_ring + rs * WIDE_BLOCK_BYTES)), blo_d = bhi_d + (uint64_t)(128 * 128 / 16);
                        const uint32_t d_tmem = tmem_base + (uint32_t)t * NT;
#pragma unroll
                        for (int pass = 0; pass < (MCR_WIDE_ABLATE == 3 ? 0 : 3); ++pass) {
                            const uint64_t ba = pass == 1 ? blo_d : bhi_d;   // transform: hi, lo, hi
                            const uint64_t za = pass == 2 ? zlo_d : zhi_d;   // normals:   hi, hi, lo
#pragma unroll
                            for (int ks = 0; ks < KBLOCK / 16; ++ks)
                                umma_f16(d_tmem, ba + 2 * ks, za + 2 * ks, IDESC, 1u);
                        }
                        umma_commit(&s_bempty[rs]);
                    }
                    __syncwarp();
                }
                if (elect_one()) {
                    umma_commit(&s_zempty[slot]);
                    if ((kb & 1) && (kb >> 1) < n_early) umma_commit(&s_slabdone[kb >> 1]);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&s_done);
            __syncwarp();
        }
    } else if (warp == LOAD_WARP) {
        // ---- transform loader: the same sequence of 32 KB blocks for every tile ----
        if (lane == 0) {
            uint32_t blk = 0;
            for (uint64_t n = 0; n < my_tiles; ++n) {
                uint32_t idx = 0;
                for (int kb = 0; kb < KBN; ++kb)
                    for (int t = p.lower ? (kb >> 1) : 0; t < NSLAB; ++t, ++blk, ++idx) {
                        const uint32_t rs = blk % WIDE_RING;
                        if (blk >= WIDE_RING) mbar_wait(&s_bempty[rs], ((blk / WIDE_RING) - 1u) & 1u);
                        if (MCR_WIDE_ABLATE == 1 && blk >= WIDE_RING) {
                            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_bfull[rs])) : "memory");
                            continue;
                        }
                        mbar_expect_tx(&s_bfull[rs], WIDE_BLOCK_BYTES);
                        bulk_g2s(smem_u32(s_ring + rs * WIDE_BLOCK_BYTES), p.image + (size_t)idx * WIDE_BLOCK_BYTES, WIDE_BLOCK_BYTES,
                                 &s_bfull[rs]);
                    }
            }
        }
        __syncwarp();
    } else {
        // ---- generator + epilogue warps ----
        const int trial = tid % NT, part = tid / NT;
        const bool active = part < PARTS;
        const uint32_t z_row = smem_u32(s_z) + (uint32_t)(trial / 8) * 1024u + (uint32_t)(trial % 8) * 128u;
        const uint32_t r7s = (uint32_t)(trial % 8) << 4;
        const int quarter = warp & 3;                        // TMEM lanes 32 * quarter .. : rows of a slab
        const uint32_t taddr_q = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int wq = warp >> 2;                            // which of the quarter's QWARPS warps: owns pieces wq, wq + QWARPS, ...
        uint32_t ukb = 0, done_parity = 0;

        // the 32 x 32 pieces (slab t, trial group g) of this quarter's lanes are dealt round-robin to its four warps
        auto init_accumulators = [&]() {
#pragma unroll 1
            for (int pc = wq; pc < NSLAB * G32; pc += QWARPS) {
                const int t = pc / G32, g = pc - t * G32;
                const uint32_t c = __float_as_uint(__ldg(p.c_scaled + t * 128 + quarter * 32 + lane));   // (mu_i + x) * 2^10 of this lane's row
                tmem_st32_bcast(taddr_q + (uint32_t)(t * NT + 32 * g), c);
            }
            tmem_st_wait();
            tc_fence_before();
            named_bar_arrive(WIDE_BAR_ACC, ALL);
        };

        // Threshold + count the accumulators of slabs [t0, t1) into the per-trial shared-memory state (s_S, s_namb, s_amb) and,
        // with reinit, hand each 32 x 32 piece back initialised for the next tile.  zsq = the thread's running sum of squares:
        // the columns generated so far cover every column these slabs have multiplied.  The caller guarantees that the MMAs
        // of the slabs are complete (mbarrier wait).
        auto slab_epilogue = [&](int t0, int t1, float zsq, bool reinit) {
            if (active) s_zsq[part][trial] = zsq;
            named_bar_sync(WIDE_BAR_GEN, WIDE_GEN_THREADS);
            tc_fence_after();
            float zn[G32], znmax = 0.0f;
#pragma unroll
            for (int g = 0; g < G32; ++g) {
                float s = 0.0f;
#pragma unroll
                for (int q = 0; q < PARTS; ++q) s += s_zsq[q][32 * g + lane];
                zn[g] = sqrtf(s) * 1.0001f;      // ||z||_2 of trial 32 g + lane over the columns so far
                znmax = fmaxf(znmax, zn[g]);
            }
            znmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(znmax)));
            uint32_t Sloc[G32];
#pragma unroll
            for (int g = 0; g < G32; ++g) Sloc[g] = 0u;
#pragma unroll 1
            for (int pc = wq; pc < (t1 - t0) * G32; pc += QWARPS) {
                const int t = t0 + pc / G32, g = pc % G32;
                const float dr = p.delta_rel[t], da = p.delta_abs[t];
                const float thr_max = __fmaf_rn(dr, znmax, da);
                uint32_t v[32];
                tmem_ld32(taddr_q + (uint32_t)(t * NT + 32 * g), v);
                tmem_ld_wait();
                // this lane's row: sign bits of its 32 trials (bit c = trial c), smallest |accumulator|
                uint32_t mlo = 0u, mhi = 0u;
                float m = 3.0e38f;
#pragma unroll
                for (int c = 15; c >= 0; --c) {
                    mlo = __funnelshift_l(v[c], mlo, 1);
                    mhi = __funnelshift_l(v[c + 16], mhi, 1);
                    m = fminf(m, fminf(fabsf(__uint_as_float(v[c])), fabsf(__uint_as_float(v[c + 16]))));
                }
                // 32 x 32 bit transpose across the warp: afterwards lane c holds trial c's sign bits over the 32 rows
                uint32_t x = (mhi << 16) | mlo;
                x = transpose_step<16, 0xFFFF0000u>(x, lane);
                x = transpose_step<8, 0xFF00FF00u>(x, lane);
                x = transpose_step<4, 0xF0F0F0F0u>(x, lane);
                x = transpose_step<2, 0xCCCCCCCCu>(x, lane);
                x = transpose_step<1, 0xAAAAAAAAu>(x, lane);
                const uint32_t hits = (uint32_t)__popc(x);
#pragma unroll
                for (int gg = 0; gg < G32; ++gg) Sloc[gg] += gg == g ? hits : 0u;
                if (__any_sync(0xffffffffu, m < thr_max)) {   // rare: look at the decisions of this piece one trial at a time
                    float zn_g = zn[0];
#pragma unroll
                    for (int gg = 1; gg < G32; ++gg) zn_g = gg == g ? zn[gg] : zn_g;
                    const float thr_mine = __fmaf_rn(dr, zn_g, da);
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float thr_c = __shfl_sync(0xffffffffu, thr_mine, c);
                        const uint32_t near = __ballot_sync(0xffffffffu, fabsf(__uint_as_float(v[c])) < thr_c);
                        if (near != 0u) {
                            const uint32_t signs = __ballot_sync(0xffffffffu, (v[c] >> 31) != 0u);
                            if (lane == 0) {
                                const uint32_t first = (uint32_t)__ffs((int)near) - 1u;
                                atomicAdd(&s_namb[32 * g + c], (uint32_t)__popc(near));
                                s_amb[32 * g + c] = ((uint32_t)(t * 128 + quarter * 32 + (int)first) << 1) | ((signs >> first) & 1u);
                            }
                        }
                    }
                }
                if (reinit)   // the piece is read: give it back holding (mu_i + x) * 2^10 of this lane's row for the next tile
                    tmem_st32_bcast(taddr_q + (uint32_t)(t * NT + 32 * g), __float_as_uint(__ldg(p.c_scaled + t * 128 + quarter * 32 + lane)));
            }
#pragma unroll
            for (int g = 0; g < G32; ++g)
                if (Sloc[g]) atomicAdd(&s_S[32 * g + lane], Sloc[g]);
            if (reinit) tmem_st_wait();
            tc_fence_before();
            named_bar_sync(WIDE_BAR_GEN, WIDE_GEN_THREADS);   // s_S / s_namb / s_amb are complete for these slabs; s_zsq may be rewritten
        };

        // the trials of a finished tile: counts, or the fix-up lists when a decision was near the threshold
        auto finish_tile = [&](uint64_t i0) {
            if (tid < NT) {
                const uint64_t i = i0 + tid;
                const uint32_t S = s_S[tid], n_amb = s_namb[tid], amb = s_amb[tid];
                s_S[tid] = 0u;
                s_namb[tid] = 0u;
                if (i < p.M) {
                    if (n_amb == 0u) {
                        cnt += (S >= (uint32_t)p.k) ? 1u : 0u;
                        sum_s += S;
                        if (p.S_out) p.S_out[i] = (int32_t)S;
                    } else if (n_amb == 1u && p.pair_list) {
                        const uint32_t slot = atomicAdd(p.pair_count, 1u);
                        if (slot < p.pair_cap)
                            p.pair_list[slot] = ((unsigned long long)i << 21) | ((unsigned long long)(S - (amb & 1u)) << 10) | (unsigned long long)(amb >> 1);
                    } else {
                        const uint32_t slot = atomicAdd(p.flag_count, 1u);
                        if (slot < p.flag_cap) p.flag_list[slot] = (uint32_t)i;
                    }
                }
            }
        };

        auto boundary_epilogue = [&](uint64_t i0, float zsq, bool last) {
            mbar_wait<64>(&s_done, done_parity);
            done_parity ^= 1u;
            slab_epilogue(n_early, NSLAB, zsq, !last);
            if (!last) named_bar_arrive(WIDE_BAR_ACC, ALL);   // every slab is initialised again (tcgen05.st waited for, fenced)
            finish_tile(i0);
        };

        if (my_tiles > 0) init_accumulators();
        float prev_zsq = 0.0f;
        uint64_t prev_i0 = 0;
        for (uint64_t n = 0; n < my_tiles; ++n) {
            const uint64_t tile = blockIdx.x + (uint64_t)gridDim.x * n;
            const uint64_t i0 = tile * NT;
            const uint64_t path = p.path_offset + i0 + (uint64_t)trial;
            PhiloxPath rng;
            rng.init((uint32_t)path, (uint32_t)(path >> 32), p.key1);
            float zsq = 0.0f;
#pragma unroll 1
            for (int kb = 0; kb < KBN; ++kb, ++ukb) {
                const uint32_t slot = ukb % ZS;
                // every thread (also the ones that generate nothing) stays within ZS K-blocks of the MMA warp
                if (ukb >= (uint32_t)ZS) mbar_wait<64>(&s_zempty[slot], ((ukb / ZS) - 1u) & 1u);
                if (ELAG > 0 && kb >= 1 + ELAG && ((kb - 1 - ELAG) & 1) == 0 && (kb - 1 - ELAG) / 2 < n_early) {
                    const int t = (kb - 1 - ELAG) / 2;
                    mbar_wait<64>(&s_slabdone[t], (uint32_t)(n & 1u));
                    slab_epilogue(t, t + 1, zsq, n + 1 < my_tiles);
                }
                if (active && !(MCR_WIDE_ABLATE == 2 && ukb >= 2)) {
                    const int col0 = KBLOCK * kb + PCOLS * part;
                    float z[PCOLS];
                    if (col0 + PCOLS <= p.N) {
                        rng.normals_multi<PCOLS / 4>((uint32_t)(col0 >> 2), p.key0, p.key1, z);
                    } else {
#pragma unroll
                        for (int q = 0; q < PCOLS; ++q) z[q] = 0.0f;
#pragma unroll
                        for (int b = 0; b < PCOLS / 4; ++b)
                            if (col0 + 4 * b < p.N) rng.normals((uint32_t)((col0 >> 2) + b), p.key0, p.key1, z + 4 * b);
#pragma unroll
                        for (int q = 0; q < PCOLS; ++q)
                            if (col0 + q >= p.N) z[q] = 0.0f;
                    }
#pragma unroll
                    for (int q = 0; q < PCOLS; ++q) zsq = fmaf(z[q], z[q], zsq);
                    const uint32_t zhi_row = z_row + slot * ZSLOT, zlo_row = zhi_row + NT * 128;
                    float unused = 0.0f;
#pragma unroll
                    for (int w = 0; w < PCOLS / 8; ++w) {
                        float zc[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) zc[q] = z[8 * w + q];
                        store_chunk(zc, part * (PCOLS / 8) + w, zhi_row, zlo_row, r7s, unused);
                    }
                }
                fence_async_smem();
                named_bar_arrive(1 + slot, ALL);
                // the previous tile's last slab(s), once K-blocks 0 .. kb_boundary of this tile are queued for the tensor core
                if (kb == kb_boundary && n > 0) boundary_epilogue(prev_i0, prev_zsq, false);
            }
            prev_i0 = i0;
            prev_zsq = zsq;
        }
        if (my_tiles > 0) boundary_epilogue(prev_i0, prev_zsq, true);
        if (tid < NT) {
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            sum_s = __reduce_add_sync(0xffffffffu, sum_s);
            if (lane == 0) { s_cnt[warp] = cnt; s_sum[warp] = sum_s; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        uint64_t c = 0, s = 0;
        for (int w = 0; w < G32; ++w) { c += s_cnt[w]; s += s_sum[w]; }
        p.block_partials[2 * (uint64_t)blockIdx.x] = my_tiles ? c : 0;
        p.block_partials[2 * (uint64_t)blockIdx.x + 1] = my_tiles ? s : 0;
    }
    if (warp == MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side -----------------------------------------------------------------------------------------
bool wide_supported(int N) { return N > 128 && N <= 128 * WIDE_MAX_SLABS; }
int wide_slabs(int N) { return (N + 127) / 128; }
int wide_kblocks(int N) { return (N + tc::KBLOCK - 1) / tc::KBLOCK; }
int wide_trials_per_tile(int N) { const int ns = wide_slabs(N); return ns <= 4 ? 128 : ns == 5 ? 96 : 64; }

static int wide_blocks(int N, bool lower) {
    int n = 0;
    for (int kb = 0; kb < wide_kblocks(N); ++kb) n += wide_slabs(N) - (lower ? kb / 2 : 0);
    return n;
}

size_t wide_image_bytes(int N, bool lower) { return (size_t)wide_blocks(N, lower) * WIDE_BLOCK_BYTES; }

// image: transform blocks (kb outer, t inner), each [hi 16 KB | lo 16 KB] of 128 rows x 64 columns of s_i * A_ij (row scales: tc_row_scale) in the
// K-major 128-byte-swizzled operand layout;  c_scaled: nslab * 128 floats
void wide_build(int N, bool lower, const double* A, const double* c, uint8_t* image, float* c_scaled, double* row_scale) {
    const int ns = wide_slabs(N), kbn = wide_kblocks(N);
    for (int i = 0; i < N; ++i) {
        double mx = std::fabs(c[i]);
        for (int k = 0; k < N; ++k) mx = std::max(mx, std::fabs(A[(size_t)i * N + k]));
        row_scale[i] = tc_row_scale(mx);
    }
    memset(image, 0, wide_image_bytes(N, lower));
    uint8_t* blk = image;
    for (int kb = 0; kb < kbn; ++kb)
        for (int t = lower ? kb / 2 : 0; t < ns; ++t, blk += WIDE_BLOCK_BYTES)
            for (int r = 0; r < 128; ++r) {
                const int i = 128 * t + r;
                if (i >= N) break;
                for (int kk = 0; kk < tc::KBLOCK; ++kk) {
                    const int j = tc::KBLOCK * kb + kk;
                    if (j >= N) break;
                    const double v = A[(size_t)i * N + j] * row_scale[i];
                    const __half hi = __float2half_rn((float)v);
                    const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
                    const uint32_t off = tc::operand_offset(128, r, kk);
                    memcpy(blk + off, &hi, 2);
                    memcpy(blk + 128 * 128 + off, &lo, 2);
                }
            }
    for (int i = 0; i < ns * 128; ++i) c_scaled[i] = (float)(i < N ? c[i] * row_scale[i] : TC_PAD_ROW);
}

template <int NT>
static cudaError_t wide_launch_nt(const WideParams& p, int num_sms, cudaStream_t s) {
    const size_t smem = (size_t)wide_zslots(NT) * (size_t)(2 * NT * 128) + (size_t)wide_ring(NT) * WIDE_BLOCK_BYTES + 1024;
    cudaError_t e = cudaFuncSetAttribute(tc_wide_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t tiles = (p.M + NT - 1) / NT;
    const int grid = (int)std::min<uint64_t>((uint64_t)num_sms, tiles ? tiles : 1);
    tc_wide_kernel<NT><<<grid, wide_gen_threads(NT) + 64, smem, s>>>(p);
    return cudaGetLastError();
}

int wide_grid(uint64_t M, int N, int num_sms) {
    const int nt = wide_trials_per_tile(N);
    const uint64_t tiles = (M + nt - 1) / nt;
    return (int)std::min<uint64_t>((uint64_t)num_sms, tiles ? tiles : 1);
}

cudaError_t wide_launch(const WideLaunch& l, int num_sms, cudaStream_t s) {
    WideParams p{};
    p.M = l.M; p.path_offset = l.path_offset; p.key0 = l.key0; p.key1 = l.key1;
    p.N = l.N; p.k = l.k; p.kbn = wide_kblocks(l.N); p.nslab = wide_slabs(l.N); p.lower = l.lower;
    p.image = reinterpret_cast<const uint8_t*>(l.d_image); p.c_scaled = l.d_c_scaled;
    for (int t = 0; t < p.nslab; ++t) {
        p.delta_rel[t] = (float)l.delta_rel[t];   // already in accumulator units (row scales applied)
        p.delta_abs[t] = (float)l.delta_abs[t];
    }
    p.S_out = l.S_out; p.flag_list = l.flag_list; p.flag_count = l.flag_count; p.flag_cap = l.flag_cap;
    p.pair_list = l.pair_list; p.pair_count = l.pair_count; p.pair_cap = l.pair_cap; p.block_partials = l.block_partials;
    switch (wide_trials_per_tile(l.N)) {
        case 128: return wide_launch_nt<128>(p, num_sms, s);
        case 96: return wide_launch_nt<96>(p, num_sms, s);
        default: return wide_launch_nt<64>(p, num_sms, s);
    }
}

}  // namespace mcr
