// Streamed tensor-core kernel for wide portfolios (128 < N <= 1024: the Stoxx600 shape; any N with supplied normals).
//
// Round 1 had two kernels here.  The single-pass kernel (mcr_wide.cu) turns the GEMM around (M = 128 transform rows, N = 96
// trials) so that all accumulators fit TMEM, and is bound by shared-memory bandwidth: every MMA re-reads a 4 KB transform
// tile, 10.4 KB of shared-memory traffic per MMA in all, 109 clocks per MMA against the nominal 48
// (profiles/r01_wide_n600_after_fix_ncu_summary.txt, r02_ubench_umma_ts.log).  The row-slab kernel (mcr_tc.cu) keeps
// M = trials but re-derives the normals once per 256-row slab.  This kernel keeps what works in each:
//
//     D[128 trials x W assets] += Z_kb[128 x 64] * B_{p,kb}[W x 64]^T            (tcgen05.mma, M = 128, N <= 256, K = 16)
//
//   * M = trials.  The A operand IS the block of normals, and it lives in TMEM (".ts" form): the generator threads write their
//     trial's fp16 hi / lo pairs straight into the tile with tcgen05.st (lane = trial; no swizzle arithmetic, no shared-memory
//     traffic, no proxy fence).  Measured: with A in TMEM an N >= 192 MMA retires at the nominal N / 2 clocks.
//   * Shared memory holds nothing but the ring of transform pieces streaming from L2 (cp.async.bulk + mbarrier, 4 x 48 KB).
//   * TMEM: 2 Z slots of 64 columns (32 hi + 32 lo) + up to 384 accumulator columns.  A model with more rows than that is done
//     in passes over row ranges; pass p only needs the normals of columns [0, kend_p), which are re-derived from the
//     counter-based stream: 224 + 600 normals per trial at N = 600 (lower-triangular), against 600 in the single-pass kernel
//     and 1368 in the row-slab kernel.  N <= 384 is a single pass.
//   * Lower-triangular transforms: K-block kb only multiplies the rows >= 64 kb (16-row granularity: the MMA's N and the
//     accumulator column offset shrink with kb), so the tensor pipe executes 1.2x the triangle instead of 1.6x.
//   * The constant mu_i + x enters as the first MMA of a pass (which therefore also initialises the accumulators): a Z slot
//     holding ones in K columns 0, 1 times a piece holding the hi and lo halves of the scaled constants.
//   * fp16 x 3 as everywhere: Zhi Bhi + Zlo Bhi + Zhi Blo, fp32 accumulate.  The hi piece of a K-block is consumed by the
//     first two products and released before the lo piece is needed.
//   * Epilogue in the same threads: lane = trial, so a thread counts the sign bits of its own row; the two smallest
//     |accumulator| keys of the row decide between "no decision near the threshold", "exactly one" (that row alone goes to
//     exact_pairs_kernel) and "several" (the whole trial goes to the fp64 kernel), exactly like the row-slab kernel.
//   * SUPPLIED = true: the normals come from the caller (M x N fp64 in HBM, serial.c:67-69 layout) instead of Philox:
//     parity mode 1 "fed identical pre-generated normals through the interface" reaches the tensor path.  The matrix is read
//     through a 2-D TMA tensor map, one box of 16 columns x 128 trials (16 KB, 128-byte swizzle) per generator part and
//     K-block, by a loader warp of its own into one or two box slots per part (cp.async.bulk.tensor + full / empty mbarriers): a thread
//     reading its own row element by element makes every warp load touch 32 lines (measured 0.55 - 0.7 TB/s at N = 600, against
//     2.6 TB/s through the boxes; N = 256: 0.85 against 4.2 - 4.5).  Rows past M and columns past N are zero-filled by the TMA unit.  (Odd N or an unaligned
//     matrix: rows are not 16-byte aligned, the element loads stay.)
#include "../../include/mcr_b200.h"
#include "mcr_internal.h"
#include "mcr_rng.cuh"
#include "mcr_tc_common.cuh"

#include <cuda.h>   // CUtensorMap (the encode function is fetched through cudaGetDriverEntryPoint: no libcuda link)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace mcr {

// tuning only (tools/devbuild_one.sh), a bit mask: 1 = the loader copies each ring slot once and then only signals it (no transform
// stream from L2), 2 = the generators skip Philox / Box-Muller (garbage normals), 4 = the MMA warp commits without issuing MMAs.
// Results are garbage; run with no_recheck.
#ifndef MCR_STREAM_ABLATE
#define MCR_STREAM_ABLATE 0
#endif

// tuning only: -DMCR_STREAM_TRACE=1 records clock64 time stamps of CTA 0 (generator warps 0 and 12, the MMA warp, the loader) into
// g_stream_trace, read back with mcr_stream_trace_dump (tools/stream_trace.py prints a timeline of one tile)
#ifndef MCR_STREAM_TRACE
#define MCR_STREAM_TRACE 0
#endif
#if MCR_STREAM_TRACE
constexpr int TRACE_ROLES = 4, TRACE_CAP = 16384;
__device__ unsigned long long g_stream_trace[TRACE_ROLES * TRACE_CAP];
__device__ unsigned int g_stream_trace_n[TRACE_ROLES];
// the event index lives in a register of the recording thread (trace_i, declared per role): recording an event is one clock read
// and one fire-and-forget store, nothing is loaded
__device__ __forceinline__ void trace_event(int role, uint32_t& i, uint32_t tag, uint32_t val) {
    if (i < TRACE_CAP) {
        g_stream_trace[role * TRACE_CAP + i] = ((unsigned long long)tag << 56) | ((unsigned long long)(val & 0xFFFFu) << 40) | ((unsigned long long)clock64() & 0xFFFFFFFFFFull);
        ++i;
        g_stream_trace_n[role] = i;
    }
}
#define ST_TRACE(role, tag, val) do { if (blockIdx.x == 0 && lane == 0) trace_event(role, trace_i, tag, (uint32_t)(val)); } while (0)
#define ST_TRACE_DECL uint32_t trace_i = 0;
#else
#define ST_TRACE(role, tag, val) do { } while (0)
#define ST_TRACE_DECL
#endif

constexpr int ST_WMAX = STREAM_WMAX;                 // accumulator columns = transform rows per pass
constexpr int ST_ZSLOTS = 2;
constexpr int ST_ZCOLS = 64;                         // TMEM columns of a Z slot: 32 (hi pairs) + 32 (lo pairs)
constexpr int ST_ACC_COL = ST_ZSLOTS * ST_ZCOLS;     // first accumulator column
constexpr int ST_CHUNK_MAX = 256;                    // rows of one transform piece = the N of the MMAs that read it
constexpr uint32_t ST_PIECE_MAX = ST_CHUNK_MAX * 128;   // 32 KB: up to 256 rows x 64 fp16
#ifndef MCR_STREAM_RING_MAX
#define MCR_STREAM_RING_MAX 6
#endif
constexpr int ST_RING_MAX = MCR_STREAM_RING_MAX;     // transform pieces in flight (6 x 32 KB; 3 when the Z cache takes its share)
constexpr int ST_RING_SLOTS = 2 * ST_RING_MAX;       // CTA-pair form: each CTA holds half of every piece, twice as many 16 KB slots
#ifndef MCR_STREAM_KCACHE_MAX
#define MCR_STREAM_KCACHE_MAX 3
#endif
constexpr int ST_KCACHE_MAX = MCR_STREAM_KCACHE_MAX;                     // K-blocks of normals kept in shared memory for the later passes (32 KB each)
constexpr uint32_t ST_ZCACHE_BYTES = 128 * 64 * 4;   // one K-block of Z: 128 trials x 64 columns x (hi + lo)
constexpr int ST_PARTS = 4;                          // generator threads per trial: 16 of a K-block's 64 columns each
constexpr int ST_GEN_THREADS = 128 * ST_PARTS;
constexpr int ST_THREADS = ST_GEN_THREADS + 64;      // + MMA-issue warp + transform loader warp
constexpr int ST_THREADS_SUPPLIED = ST_THREADS + 32; // + Z loader warp (the supplied-normals instantiation only)
constexpr int ST_MMA_WARP = ST_GEN_THREADS / 32, ST_LOAD_WARP = ST_MMA_WARP + 1, ST_ZLOAD_WARP = ST_MMA_WARP + 2;
// supplied normals: TMA boxes (16 columns x 128 trials of fp64, 16 KB each), one or two per generator part.  Every box slot belongs to
// ONE part: its four warps take the slot's uses in order, so a parity wait is never more than one phase ahead (boxes land out of
// order; with slots shared between parts a fast warp could wait for use u + 1 of a slot whose use u has not landed yet, and a
// parity wait two phases ahead passes at once).  Shared memory is split between the boxes and the ring of transform pieces:
// single-pass models (N <= 384) are bound by the stream of normals - two boxes per part, two pieces (N = 256: 4.5 against 4.1 TB/s);
// models of several passes also need their transform stream - one box per part, four pieces (N = 600: 2.6 against 2.3 TB/s).
constexpr int ST_ZBOX_MAX = ST_PARTS * 2;
constexpr uint32_t ST_ZBOX_BYTES = 16 * 128 * 8;
constexpr int ST_ALL = ST_GEN_THREADS + 32;          // generators + MMA warp
constexpr int ST_BAR_SLOT = 1, ST_BAR_ACC = 3, ST_BAR_GEN = 4;   // named barriers: 1, 2 = Z slot full; accumulators free; generators only
constexpr uint32_t ST_KEY_MASK = 0xFFFFFC00u;        // epilogue keys: bits 10.. = |accumulator| << 1 (truncated), 1..9 = column, 0 = sign
static_assert(ST_ACC_COL + ST_WMAX <= 512, "TMEM columns");
static_assert(ST_RING_MAX * ST_PIECE_MAX + 1024 <= 220 * 1024 && (ST_RING_MAX - ST_KCACHE_MAX) * ST_PIECE_MAX + ST_KCACHE_MAX * ST_ZCACHE_BYTES + 1024 <= 220 * 1024,
              "shared memory");

struct StreamParams {
    uint64_t M, path_offset;
    uint32_t key0, key1;
    PhiloxKeys keys;           // the round keys of (key0, key1): constant-bank operands of the generator
    int32_t N, k, lower, npass;
    int32_t ring;              // ring slots in use
    int32_t kcache;            // K-blocks 0 .. kcache - 1 of a tile's normals are kept in shared memory after the first pass
    StreamPass pass[STREAM_MAX_PASSES];
    float delta_rel[STREAM_MAX_PASSES], delta_abs[STREAM_MAX_PASSES], thr_max[STREAM_MAX_PASSES];   // accumulator units
    const uint8_t* image;      // pieces in consumption order (stream_build)
    size_t image_rank_bytes;   // CTA-pair form: the image is [halves of CTA 0][halves of CTA 1], this many bytes each
    const double* Z;           // SUPPLIED: M x N fp64 row-major (chunk-relative), else nullptr
    int32_t ztma;              // SUPPLIED: 1 = Z is read through the tensor map (second kernel parameter), 0 = element loads
    int32_t zbox_shift;        // ... into 1 << zbox_shift box slots (4 or 8: one or two per generator part)
    int32_t* S_out;
    uint32_t* flag_list;
    uint32_t* flag_count;
    uint32_t flag_cap;
    unsigned long long* pair_list;   // trial << 21 | other hits << 10 | row, or nullptr
    uint32_t* pair_count;
    uint32_t pair_cap;
    uint64_t* block_partials;
    // portfolio extension (north-star A7; no reference code): accumulator column port_col of the LAST pass holds s_p (w.R)
    int32_t portfolio, port_col;
    float port_thresh;            // scaled: tail event is  D < port_thresh  (= -x_p s_p)
    float port_inv_scale;         // 1 / s_p
    float hist_lo, hist_inv_w;    // unscaled histogram origin and 1 / bin width
    double hist_lo_d, hist_w_d;   // the same origin and the bin width in fp64
    double* port_acc;             // [4] += hits, sum P, sum P^2, sum of -P over tail events
    unsigned long long* port_hist;  // [MCR_PORTFOLIO_BINS] counts
    double* port_hist_sum;          // [MCR_PORTFOLIO_BINS] per-bin sum of P
};

// first accumulator column K-block kb multiplies in a pass starting at row r0 (rows < 64 kb of a lower-triangular transform are
// zero in these columns); r0 and 64 kb are multiples of 16, so the MMA's N stays a multiple of 16
__host__ __device__ inline int stream_acol0(int r0, int kb, int lower) {
    const int a = lower ? 64 * kb - r0 : 0;
    return a > 0 ? a : 0;
}
// Two build options that were measured and are OFF (profiles/r02_stream_kernel_development.log, "stream8"):
//  * MCR_STREAM_DEFER=1: the tcgen05.st of a step are not waited for on the spot; the wait and the hand-over of the Z slot are retired
//    in the middle of the next step's generation (after its Philox rounds).  Hides the store latency, but the MMAs of every step
//    start later and the volatile wait splits the generator's instruction schedule: 2 % slower at N = 600, 10 % at N = 256.
//  * MCR_STREAM_PREFETCH=1: the last pass generates its first uncached K-block right after K-block 0 has been fed (into the cache slot
//    of K-block 0, which is dead by then), while the tensor core works through the cached K-blocks.  Neutral (+-0.5 %).
#ifndef MCR_STREAM_DEFER
#define MCR_STREAM_DEFER 0
#endif
#ifndef MCR_STREAM_PREFETCH
#define MCR_STREAM_PREFETCH 0
#endif
// Order in which a pass visits its K-blocks: ascending by default.  MCR_STREAM_CACHED_LAST=1 (build option) makes a later pass take
// the K-blocks it has to generate first (kcache, kcache + 1, ...) and the cached ones (0 .. kcache - 1) last - feeding a cached block
// costs the generators nothing while its MMAs are the widest of the pass, so at the end of the pass the generators could already
// produce the next tile.  Measured: 2 % SLOWER at N = 600 / 1000 (profiles/r02_stream_kernel_development.log), because the
// incremental read-out of finished columns is lost in those passes; kept for A/B runs.
#ifndef MCR_STREAM_CACHED_LAST
#define MCR_STREAM_CACHED_LAST 0
#endif
__host__ __device__ inline int stream_kb(int ps, int step, int kbn, int kcache) {
    if (!MCR_STREAM_CACHED_LAST || ps == 0 || kcache <= 0) return step;
    return step < kbn - kcache ? kcache + step : step - (kbn - kcache);
}
__host__ __device__ inline bool stream_ascending(int ps, int kcache) { return !MCR_STREAM_CACHED_LAST || ps == 0 || kcache <= 0; }

// an active width w <= 384 as one or two chunks (one transform piece and one MMA N each; <= 256, multiple of 16)
__host__ __device__ inline int stream_n1(int w) { return w <= ST_CHUNK_MAX ? w : ((w / 2 + 15) & ~15); }

namespace tc {
// D[tmem] (+)= A[tmem] * B[smem desc]^T: the A operand (128 lanes x 8 columns of fp16 pairs per K-step) comes from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_store_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx_st(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_st(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the same copy delivered to the same shared-memory offset (and signalled on the mbarrier at the same offset) of every CTA in cta_mask
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
// tcgen05.commit arriving on the mbarrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
// one box of a 2-D tensor map (column c0, row r0) into shared memory, signalled on the mbarrier with its bytes
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int32_t c0, int32_t r0, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
                 "l"(map), "r"(c0), "r"(r0), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// ---- CTA-pair form (tcgen05 cta_group::2; layout and hand-shake checked by tools/pair_layout_check.cu) ----
// one MMA over both CTAs of the pair: M = 256 (each CTA's 128 A rows from its own TMEM), B = [rows of CTA 0 ; rows of CTA 1] from
// the same shared-memory offset in both, the accumulator at the same TMEM columns in both; issued by the leader (rank 0) alone
__device__ __forceinline__ void umma_f16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// arrive (release, cluster scope) on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
// wait on a barrier that threads of the other CTA arrive on (acquire at cluster scope)
// (backs off between failed polls: the MMA warp shares its scheduler with four generator warps)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "WAIT_LOOP:\n"
        "nanosleep.u32 %2;\n"
        "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(20u) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_pair(int n) {   // as make_idesc, M = 256
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
}  // namespace tc

// 16 normals -> 8 + 8 packed fp16 pairs (hi, lo); same split as store_chunk: hi = rn16(z), lo = rn16(z - hi)
__device__ __forceinline__ void split16(const float (&z)[16], uint32_t (&hv)[8], uint32_t (&lv)[8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const __half2 h = __floats2half2_rn(z[2 * q], z[2 * q + 1]);
        hv[q] = *reinterpret_cast<const uint32_t*>(&h);
        const __half2 l = __floats2half2_rn(sub_half(z[2 * q], (unsigned short)(hv[q] & 0xffffu)), sub_half(z[2 * q + 1], (unsigned short)(hv[q] >> 16)));
        lv[q] = *reinterpret_cast<const uint32_t*>(&l);
    }
}

// PORT: the image carries the portfolio row (a separate instantiation: the plain kernel pays nothing for the extension)
// PAIR: the two CTAs of a cluster work as one tensor-core pair (cta_group::2).  Each CTA still generates and reads out its own
// tile of 128 trials, but it fetches only HALF of every transform piece (rows [r n / 2, (r + 1) n / 2) of the chunk in CTA r) and
// the leader's MMA warp issues one M = 256 MMA for both tiles: per SM, half the bytes taken in from L2 and half the shared-memory
// operand reads.  The hand-overs that are named barriers in the single-CTA form (Z slot written, accumulators read out) become
// cluster-scope mbarriers on the leader; the MMA warp of CTA 1 only relays "my half of the piece has arrived" to the leader.
template <bool SUPPLIED, bool PORT, bool PAIR>
__global__ void __launch_bounds__(SUPPLIED ? ST_THREADS_SUPPLIED : ST_THREADS, 1) tc_stream_kernel(const StreamParams p, const __grid_constant__ CUtensorMap zmap) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* s_ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // [ring][PIECE]
    const int RING = p.ring;
    constexpr uint32_t PIECE = PAIR ? ST_PIECE_MAX / 2 : ST_PIECE_MAX;   // bytes of a ring slot
    uint4* s_zcache = reinterpret_cast<uint4*>(s_ring + (size_t)RING * PIECE);   // [kcache][part][4][128 rows] x 16 B
    // portfolio histogram of this CTA (only with p.portfolio): count << 44 | quantised position sum per bin (port_hist_word)
    unsigned long long* s_hist = reinterpret_cast<unsigned long long*>(s_ring + (size_t)RING * PIECE + (size_t)p.kcache * ST_ZCACHE_BYTES);
    __shared__ uint64_t s_zempty[ST_ZSLOTS], s_bfull[ST_RING_SLOTS], s_bempty[ST_RING_SLOTS], s_done;
    // PAIR, used in the leader only: Z slot written / accumulators read out by the generators of both CTAs; CTA 1's half piece arrived
    __shared__ uint64_t s_zfull[ST_ZSLOTS], s_accfree, s_bpeer[ST_RING_SLOTS];
    // SUPPLIED through TMA: the ring of Z boxes behind the transform ring; filled by the Z loader, each read by the four warps of a part
    __shared__ uint64_t s_zbfull[ST_ZBOX_MAX], s_zbempty[ST_ZBOX_MAX];
    uint8_t* const s_zbox = s_ring + (size_t)RING * PIECE;
    const bool ZTMA = SUPPLIED && p.ztma != 0;
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_part_S[ST_PARTS][128], s_part_k1[ST_PARTS][128], s_part_k2[ST_PARTS][128];
    __shared__ float s_part_zsq[ST_PARTS][128];
    __shared__ uint32_t s_cnt[4], s_sum[4];

    // the warp index through a shuffle: warp-uniform for the compiler (role loops keep their state in uniform registers)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    // CTA pairs (clusters of 2, the two SMs of a TPC) share the transform stream: each CTA fetches every other piece and multicasts
    // it into both rings.  At full tensor rate one SM consumes ~43 bytes of transform per clock, which is the whole chip's L2
    // bandwidth divided by 148 (profiles/r02_stream_kernel_development.log): unshared, the L2 stream and the MMAs take turns.
    const uint32_t csz = cluster_nctarank(), crank = cluster_ctarank();
    const uint16_t cmask = (uint16_t)((1u << csz) - 1u);
    if (tid == 0) {
        for (int b = 0; b < ST_ZSLOTS; ++b) mbar_init(&s_zempty[b], 1);
        // (PAIR: the leader's one commit arrives once on each CTA's copy; otherwise every CTA of the cluster commits to all copies)
        for (int b = 0; b < ST_RING_SLOTS; ++b) { mbar_init(&s_bfull[b], 1); mbar_init(&s_bempty[b], PAIR ? 1u : csz); mbar_init(&s_bpeer[b], 1); }
        mbar_init(&s_done, 1);
        for (int b = 0; b < ST_ZSLOTS; ++b) mbar_init(&s_zfull[b], 2 * ST_GEN_THREADS / 32);   // one arrival per generator warp of the pair
        mbar_init(&s_accfree, 2 * ST_GEN_THREADS / 32);
        if (SUPPLIED)
            for (int b = 0; b < ST_ZBOX_MAX; ++b) { mbar_init(&s_zbfull[b], 1); mbar_init(&s_zbempty[b], 4); }
    }
    if (PORT)
        for (int t = tid; t < MCR_PORTFOLIO_BINS; t += ST_THREADS) s_hist[t] = 0ull;
    if (!PAIR && warp == ST_MMA_WARP) tmem_alloc(&s_tmem, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (csz > 1) cluster_sync_all();   // the partner's barriers exist before anything is multicast to them
    if (PAIR) {                        // both CTAs allocate the same columns, together
        if (warp == ST_MMA_WARP) tmem_alloc_pair(&s_tmem, 512);
        tc_fence_before();
        __syncthreads();
        cluster_sync_all();
    }
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    // every CTA runs the same number of tiles (a pair must consume the shared stream in step); tiles past the end are ghosts
    // whose trials (i >= M) are generated and multiplied like any other but never counted
    const uint64_t n_tiles = (p.M + TILE_M - 1) / TILE_M;
    const uint64_t my_tiles = (n_tiles + gridDim.x - 1) / gridDim.x;
    ST_TRACE_DECL
    const int NPASS = p.npass, LOWER = p.lower;

    if (warp == ST_MMA_WARP) {
        // ---- MMA issue: per pass the constant slot, then the K-blocks in order.  A K-block arrives as one hi and one lo piece per
        // chunk of <= 256 rows; a hi piece is multiplied by Zhi and Zlo, a lo piece by Zhi, and each is released right after ----
        uint32_t zu = 0, rs = 0, rphase = 0, passes = 0;
        auto next_piece = [&]() { if (++rs == (uint32_t)RING) { rs = 0; rphase ^= 1u; } };
        if (PAIR && crank != 0) {
            // CTA 1 of a pair: tell the leader when this CTA's half of each piece has landed
            for (uint64_t n = 0; n < my_tiles; ++n)
                for (int ps = 0; ps < NPASS; ++ps) {
                    const int r0 = p.pass[ps].r0, W = p.pass[ps].W, KBN = p.pass[ps].kbn;
                    for (int st = -1; st < KBN; ++st) {
                        const int kb = st < 0 ? -1 : stream_kb(ps, st, KBN, p.kcache);
                        const int w = W - (kb < 0 ? 0 : stream_acol0(r0, kb, LOWER));
                        const int npieces = (kb < 0 ? 1 : 2) * (w > ST_CHUNK_MAX ? 2 : 1);
                        for (int q = 0; q < npieces; ++q) {
                            mbar_wait<20>(&s_bfull[rs], rphase);
                            if (lane == 0) mbar_arrive_remote(&s_bpeer[rs], 0);
                            __syncwarp();
                            next_piece();
                        }
                    }
                }
        } else
        for (uint64_t n = 0; n < my_tiles; ++n) {
            for (int ps = 0; ps < NPASS; ++ps, ++passes) {
                const int r0 = p.pass[ps].r0, W = p.pass[ps].W, KBN = p.pass[ps].kbn;
                for (int st = -1; st < KBN; ++st, ++zu) {
                    const int kb = st < 0 ? -1 : stream_kb(ps, st, KBN, p.kcache);
                    const uint32_t slot = zu & 1u;
                    ST_TRACE(2, 1, kb);                                    // waiting for the Z slot
                    // the Z slot is written (tcgen05.st waited for, fenced) - in both CTAs of a pair
                    if (PAIR) mbar_wait_cluster(&s_zfull[slot], (zu >> 1) & 1u);
                    else named_bar_sync(ST_BAR_SLOT + slot, ST_ALL);
                    ST_TRACE(2, 2, kb);
                    if (kb < 0) {                                          // the previous pass has been read out of the accumulators
                        if (PAIR) mbar_wait_cluster(&s_accfree, passes & 1u);
                        else named_bar_sync(ST_BAR_ACC, ST_ALL);
                        ST_TRACE(2, 3, kb);
                    }
                    const int acol0 = kb < 0 ? 0 : stream_acol0(r0, kb, LOWER);
                    const int w = W - acol0, n1 = stream_n1(w), n2 = w - n1;
                    const uint32_t a_hi = tmem_base + slot * ST_ZCOLS, a_lo = a_hi + 32u;
                    const uint32_t d1 = tmem_base + (uint32_t)(ST_ACC_COL + acol0);
                    for (int half = 0; half < (kb < 0 ? 1 : 2); ++half) {
                        for (int c = 0; c < (n2 > 0 ? 2 : 1); ++c) {
                            const uint32_t d = c ? d1 + (uint32_t)n1 : d1;
                            const uint32_t id = PAIR ? make_idesc_pair(c ? n2 : n1) : make_idesc(c ? n2 : n1);
                            const bool last = half == (kb < 0 ? 0 : 1) && c == (n2 > 0 ? 1 : 0);
                            mbar_wait<20>(&s_bfull[rs], rphase);
                            if (PAIR) mbar_wait_cluster(&s_bpeer[rs], rphase);   // ... and CTA 1's half
                            ST_TRACE(2, 4, kb);                            // piece arrived
                            if (elect_one()) {
                                tc_fence_after();
                                const uint64_t b = make_desc(smem_u32(s_ring + rs * PIECE));
                                auto mma = [&](uint32_t a, uint64_t bd, uint32_t accumulate) {
                                    if (PAIR) umma_f16_ts_pair(d, a, bd, id, accumulate);
                                    else umma_f16_ts(d, a, bd, id, accumulate);
                                };
                                if (!(MCR_STREAM_ABLATE & 4)) {
                                    if (kb < 0) {   // constants: ones (K columns 0, 1) x [hi | lo] of (mu_i + x) s_i; initialises the accumulators
                                        mma(a_hi, b, 0u);
                                    } else if (half == 0) {   // hi piece: Zhi Bhi, then Zlo Bhi
#pragma unroll
                                        for (int ks = 0; ks < KBLOCK / 16; ++ks) mma(a_hi + 8 * ks, b + 2 * ks, 1u);
#pragma unroll
                                        for (int ks = 0; ks < KBLOCK / 16; ++ks) mma(a_lo + 8 * ks, b + 2 * ks, 1u);
                                    } else {                  // lo piece: Zhi Blo
#pragma unroll
                                        for (int ks = 0; ks < KBLOCK / 16; ++ks) mma(a_hi + 8 * ks, b + 2 * ks, 1u);
                                    }
                                }
                                if (PAIR) {   // every commit of the pair arrives in both CTAs
                                    umma_commit_pair(&s_bempty[rs]);
                                    if (last) {
                                        umma_commit_pair(&s_zempty[slot]);
                                        if (st == KBN - 1) umma_commit_pair(&s_done);
                                    }
                                } else {
                                    if (csz > 1) umma_commit_multicast(&s_bempty[rs], cmask);   // both rings' slot rs: one more consumer done
                                    else umma_commit(&s_bempty[rs]);
                                    if (last) {
                                        umma_commit(&s_zempty[slot]);
                                        if (st == KBN - 1) umma_commit(&s_done);
                                    }
                                }
                            }
                            __syncwarp();
                            ST_TRACE(2, 5, kb);                            // MMAs of the piece issued
                            next_piece();
                        }
                    }
                }
            }
        }
    } else if (warp == ST_LOAD_WARP) {
        // ---- transform loader: the same sequence of pieces for every tile ----
        if (lane == 0) {
            uint32_t rs = 0, rphase = 0, piece = 0;
            for (uint64_t n = 0; n < my_tiles; ++n) {
                const uint8_t* src = p.image + (PAIR ? (size_t)crank * p.image_rank_bytes : 0);   // PAIR: this CTA's halves
                for (int ps = 0; ps < NPASS; ++ps) {
                    const int r0 = p.pass[ps].r0, W = p.pass[ps].W, KBN = p.pass[ps].kbn;
                    for (int st = -1; st < KBN; ++st) {
                        const int kb = st < 0 ? -1 : stream_kb(ps, st, KBN, p.kcache);
                        const int w = W - (kb < 0 ? 0 : stream_acol0(r0, kb, LOWER)), n1 = stream_n1(w), n2 = w - n1;
                        for (int half = 0; half < (kb < 0 ? 1 : 2); ++half)
                            for (int c = 0; c < (n2 > 0 ? 2 : 1); ++c, ++piece) {
                                const uint32_t bytes = (uint32_t)(c ? n2 : n1) * (PAIR ? 64u : 128u);
                                // slot rs has been consumed by every CTA of the pair (their commits arrive on both copies of s_bempty)
                                if (piece >= (uint32_t)RING) mbar_wait<20>(&s_bempty[rs], rphase ^ 1u);
                                ST_TRACE(3, 1, kb);                        // ring slot free, copy issued
                                if ((MCR_STREAM_ABLATE & 1) && piece >= (uint32_t)RING) {
                                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_bfull[rs])) : "memory");
                                } else {
                                    mbar_expect_tx_st(&s_bfull[rs], bytes);   // this CTA's copy, whoever fetches it
                                    if (PAIR || csz == 1) bulk_g2s_st(smem_u32(s_ring + rs * PIECE), src, bytes, &s_bfull[rs]);
                                    else if ((piece & 1u) == crank) bulk_g2s_multicast(smem_u32(s_ring + rs * PIECE), src, bytes, &s_bfull[rs], cmask);
                                }
                                src += bytes;
                                if (++rs == (uint32_t)RING) { rs = 0; rphase ^= 1u; }
                            }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == ST_ZLOAD_WARP) {
        // ---- supplied normals: the Z boxes in the order the generators consume them (tile, pass, K-block, part) ----
        if (ZTMA && lane == 0) {
            uint32_t seq = 0;
            for (uint64_t n = 0; n < my_tiles; ++n) {
                const uint64_t tile = blockIdx.x + (uint64_t)gridDim.x * n;
                const int32_t r0 = (int32_t)min(tile * TILE_M, (uint64_t)0x7fffff00u);   // (a ghost tile lies past M: zero-filled)
                for (int ps = 0; ps < NPASS; ++ps)
                    for (int kb = 0; kb < p.pass[ps].kbn; ++kb)
                        for (int part = 0; part < ST_PARTS; ++part, ++seq) {
                            const uint32_t slot = seq & ((1u << p.zbox_shift) - 1u), use = seq >> p.zbox_shift;   // (slot % ST_PARTS == part)
                            if (use > 0) mbar_wait<20>(&s_zbempty[slot], (use - 1u) & 1u);
                            mbar_expect_tx_st(&s_zbfull[slot], ST_ZBOX_BYTES);
                            tma_load_2d(smem_u32(s_zbox + (size_t)slot * ST_ZBOX_BYTES), &zmap, KBLOCK * kb + 16 * part, r0, &s_zbfull[slot]);
                        }
            }
        }
        __syncwarp();
    } else {
        // ---- generator + epilogue warps: warp w owns TMEM lanes 32 (w % 4) .. +31 (trial rows) and columns 16 (w / 4) .. +15 of a K-block ----
        const int quarter = warp & 3, part = warp >> 2, row = 32 * quarter + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * quarter) << 16);
        const int KCACHE = SUPPLIED ? 0 : p.kcache;
        uint32_t zu = 0, done_parity = 0, cnt = 0, sum_s = 0;
        uint32_t zk = 0;   // (supplied normals through TMA) K-block steps taken so far: box sequence number = 4 zk + part
        // state of the trial in this thread's row across the passes of a tile (kept by the part-0 thread of the row)
        uint32_t S_tot = 0, n_amb = 0, amb = 0;

        // Read-out of the current pass, per thread: the 32-column groups g = part, part + 4, ... of the row (dealt round-robin to
        // its four threads).  c_next = first column of this thread's next group; S_pass / k1 <= k2 = the sign bits counted and the
        // two smallest keys (|accumulator|, column, sign) seen so far.
        uint32_t S_pass = 0, k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
        int c_next = 32 * part;
        // portfolio extension: partial sums of the trials whose portfolio column this thread reads
        uint32_t p_hits = 0;
        float p_sum = 0.0f, p_sum2 = 0.0f, p_tail = 0.0f;

        // "this warp has read the accumulators of the pass out": to the MMA warp (PAIR: the leader's, from both CTAs)
        auto acc_free = [&]() {
            if (PAIR) { __syncwarp(); if (lane == 0) mbar_arrive_remote(&s_accfree, 0); }
            else named_bar_arrive(ST_BAR_ACC, ST_ALL);
        };
        // read this thread's groups that end at or below column `upto` (everything left with final_part)
        auto read_columns = [&](int ps, int upto, bool final_part, uint64_t i) {
            const int W = p.pass[ps].W;
            const float thr_max = p.thr_max[ps];
#pragma unroll 1
            while (final_part ? c_next < W : c_next + 32 <= upto) {
                uint32_t v[32];
                tmem_ld32(lane_addr + (uint32_t)(ST_ACC_COL + c_next), v);
                tmem_ld_wait();
                const int valid = W - c_next;    // a multiple of 16: the last group of a pass may hold only 16 columns
                if (PORT && ps == NPASS - 1 && c_next <= p.port_col && p.port_col < c_next + 32) {
                    // this group holds the portfolio column (row N of the image: w^T A and w.mu): take P out and leave a harmless
                    // large positive value in its place for the crash count and the band
                    const int pcl = p.port_col - c_next;
                    uint32_t pbits = 0u;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (c == pcl) { pbits = v[c]; v[c] = 0x7F000000u; }
                    if (i < p.M) {
                        const float Ps = __uint_as_float(pbits), P = Ps * p.port_inv_scale;
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
                float m = 3.0e38f;
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    if (c < 16 || valid > 16) {
                        S_pass += (v[c] >> 31) + (v[c + 1] >> 31);
                        m = fminf(m, fminf(fabsf(__uint_as_float(v[c])), fabsf(__uint_as_float(v[c + 1]))));
                    }
                }
                if (m < thr_max) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (c < 16 || valid > 16) {
                            const uint32_t key = (__funnelshift_l(v[c], v[c], 1) & (ST_KEY_MASK | 1u)) | ((uint32_t)(c_next + c) << 1);
                            k2 = min(k2, max(k1, key));
                            k1 = min(k1, key);
                        }
                    }
                }
                c_next += 32 * ST_PARTS;
            }
        };

        // End of a pass: the rest of its columns, then the four threads of a row combine in shared memory and the part-0 thread
        // decides about the near-threshold decisions.
        auto epilogue = [&](int ps, float zsq, uint64_t i, bool last_of_tile, bool more_work) {
            mbar_wait<64>(&s_done, done_parity);
            done_parity ^= 1u;
            tc_fence_after();
            if (warp == 0) ST_TRACE(0, 7, ps);                                                        // the pass's MMAs are complete
            const int r0 = p.pass[ps].r0;
            read_columns(ps, p.pass[ps].W, true, i);
            uint32_t S = S_pass;
            const uint32_t k1_ = k1, k2_ = k2;
            S_pass = 0u; k1 = 0xFFFFFFFFu; k2 = 0xFFFFFFFFu; c_next = 32 * part;
            tc_fence_before();
            if (more_work) acc_free();   // the MMA warp may overwrite the accumulators (next pass)
            s_part_S[part][row] = S;
            s_part_k1[part][row] = k1_;
            s_part_k2[part][row] = k2_;
            s_part_zsq[part][row] = zsq;
            named_bar_sync(ST_BAR_GEN, ST_GEN_THREADS);
            if (part == 0) {
                uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;
                float zz = 0.0f;
                S = 0;
#pragma unroll
                for (int q = 0; q < ST_PARTS; ++q) {
                    S += s_part_S[q][row];
                    zz += s_part_zsq[q][row];
                    const uint32_t a = s_part_k1[q][row], b = s_part_k2[q][row];
                    m2 = min(m2, max(m1, a)); m1 = min(m1, a);
                    m2 = min(m2, max(m1, b)); m1 = min(m1, b);
                }
                const float thr = __fmaf_rn(p.delta_rel[ps], sqrtf(zz) * 1.0001f, p.delta_abs[ps]);
                const uint32_t thr_key = __float_as_uint(thr) << 1;
                if ((m1 & ST_KEY_MASK) < thr_key) {
                    if ((m2 & ST_KEY_MASK) < thr_key) {
                        n_amb += 2u;
                    } else {
                        if (n_amb == 0u) amb = ((uint32_t)(r0 + (int)((m1 >> 1) & 0x1FFu)) << 1) | (m1 & 1u);
                        n_amb += 1u;
                        S -= m1 & 1u;        // the uncertain decision is left out of S; exact_pairs_kernel adds the fp64 one
                    }
                }
                S_tot += S;
                if (last_of_tile) {
                    if (i < p.M) {
                        if (n_amb == 0u) {
                            cnt += (S_tot >= (uint32_t)p.k) ? 1u : 0u;
                            sum_s += S_tot;
                            if (p.S_out) p.S_out[i] = (int32_t)S_tot;
                        } else if (n_amb == 1u && p.pair_list) {
                            const uint32_t slot = atomicAdd(p.pair_count, 1u);
                            if (slot < p.pair_cap)
                                p.pair_list[slot] = ((unsigned long long)i << 21) | ((unsigned long long)S_tot << 10) | (unsigned long long)(amb >> 1);
                        } else {
                            const uint32_t slot = atomicAdd(p.flag_count, 1u);
                            if (slot < p.flag_cap) p.flag_list[slot] = (uint32_t)i;
                        }
                    }
                    S_tot = 0u; n_amb = 0u; amb = 0u;
                }
            }
            named_bar_sync(ST_BAR_GEN, ST_GEN_THREADS);   // the exchange buffers may be rewritten
        };

        if (my_tiles > 0) acc_free();   // nothing to read before the very first pass
        // Hand-over of a written Z slot to the MMA warp: wait for the tcgen05.st, fence, arrive.  (MCR_STREAM_DEFER=1 retires it in
        // the middle of the NEXT step's generation instead, or right away where the next step generates nothing.)
        bool deferred = false;
        uint32_t def_slot = 0;
        auto flush = [&]() {
            if (deferred) {
                tmem_store_wait();
                tc_fence_before();
                if (PAIR) { __syncwarp(); if (lane == 0) mbar_arrive_remote(&s_zfull[def_slot], 0); }
                else named_bar_arrive(ST_BAR_SLOT + def_slot, ST_ALL);
                deferred = false;
            }
        };
        // this thread's 16 normals of K-block kb as fp16 hi / lo pairs; zsq += their squares
        auto gen16 = [&](int kb, int kend, uint64_t path, const PhiloxPath& rng, uint32_t (&hv)[8], uint32_t (&lv)[8], float& zsq) {
            const int col0 = KBLOCK * kb + 16 * part;
            float z[16];
            if (MCR_STREAM_ABLATE & 2) {
                flush();
#pragma unroll
                for (int q = 0; q < 16; ++q) z[q] = __uint_as_float((((uint32_t)path * 747796405u + (uint32_t)(col0 + q) * 2891336453u) >> 9) | 0x3f000000u);
            } else if (col0 + 16 <= kend) {
                rng.normals_multi_between<4>((uint32_t)(col0 >> 2), p.keys, z, [&] { flush(); });
            } else {   // the K-block that holds the pass's last normal column
                flush();
#pragma unroll
                for (int q = 0; q < 16; ++q) z[q] = 0.0f;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (col0 + 4 * b < kend) rng.normals((uint32_t)((col0 >> 2) + b), p.keys, z + 4 * b);
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (col0 + q >= kend) z[q] = 0.0f;
            }
            float2 acc2 = make_float2(zsq, 0.0f);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc2 = __ffma2_rn(make_float2(z[2 * q], z[2 * q + 1]), make_float2(z[2 * q], z[2 * q + 1]), acc2);
            zsq = acc2.x + acc2.y;
            split16(z, hv, lv);
        };
        bool pending = false;
        int pend_ps = 0;
        float pend_zsq = 0.0f;
        uint64_t pend_i = 0;
        // this thread's 64 bytes per cached K-block: four 16-byte words, consecutive rows consecutive (conflict-free); written in the
        // first pass and read back by the same thread in the later ones, so the cache needs no synchronisation at all
        uint4* zc = s_zcache + (size_t)part * 4 * 128 + row;
        for (uint64_t n = 0; n < my_tiles; ++n) {
            const uint64_t tile = blockIdx.x + (uint64_t)gridDim.x * n;
            const uint64_t i = tile * TILE_M + (uint64_t)row;
            const uint64_t path = p.path_offset + i;
            PhiloxPath rng;
            if (!SUPPLIED) rng.init((uint32_t)path, (uint32_t)(path >> 32), p.key1);
            const double* zrow = SUPPLIED ? p.Z + (size_t)i * (size_t)p.N : nullptr;
            float zsq_cached = 0.0f;   // sum of squares of this thread's normals in the cached K-blocks
            for (int ps = 0; ps < NPASS; ++ps) {
                const int KBN = p.pass[ps].kbn, kend = p.pass[ps].kend;
                float zsq = ps > 0 ? zsq_cached : 0.0f;
                const bool prefetch = MCR_STREAM_PREFETCH && !MCR_STREAM_CACHED_LAST && !SUPPLIED && ps > 0 && ps == NPASS - 1 && KCACHE > 0 && KBN > KCACHE;
#pragma unroll 1
                for (int st = -1; st < KBN; ++st, ++zu) {
                    const int kb = st < 0 ? -1 : stream_kb(ps, st, KBN, KCACHE);
                    const uint32_t slot = zu & 1u;
                    const uint32_t zaddr = lane_addr + slot * ST_ZCOLS + (uint32_t)(8 * part);
                    if (warp == 0) ST_TRACE(0, 1, kb); else if (warp == 12) ST_TRACE(1, 1, kb);      // step begins
                    uint32_t hv[8], lv[8];
                    if (kb >= 0) {
                        const int col0 = KBLOCK * kb + 16 * part;
                        uint4* zck = zc + (size_t)kb * (ST_ZCACHE_BYTES / 16);
                        if (SUPPLIED) {
                            flush();
                            double zbuf[16];
                            if (ZTMA) {
                                // this part's box: row = trial, 128 bytes = 16 doubles, 16-byte chunk c stored at c ^ (row % 8)
                                const uint32_t seq = 4u * zk + (uint32_t)part, slot = seq & ((1u << p.zbox_shift) - 1u);
                                mbar_wait<20>(&s_zbfull[slot], (seq >> p.zbox_shift) & 1u);
                                const uint8_t* brow = s_zbox + (size_t)slot * ST_ZBOX_BYTES + (size_t)row * 128;
#pragma unroll
                                for (int c = 0; c < 8; ++c) {
                                    const double2 v = *reinterpret_cast<const double2*>(brow + ((c ^ (row & 7)) << 4));
                                    zbuf[2 * c] = v.x; zbuf[2 * c + 1] = v.y;
                                }
                                // The box goes back to the loader, whose refill is an ASYNC-proxy write: a proxy fence orders this
                                // thread's (generic-proxy) reads of the box before it.  Without the fence the arrive overtakes the
                                // loads (SASS: eight LDS.128, then SYNCS.ARRIVE at once) and, with the tensor core streaming its
                                // operand out of shared memory, the refill lands before they are served - measured: up to 4 % wrong
                                // trials whenever a box lay in the first 128 KB of the dynamic region, none with all boxes above
                                // (profiles/r02_stream_ztma_experiment.log).
                                fence_async_smem();
                                __syncwarp();
                                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_zbempty[slot])) : "memory");
                                ++zk;
                            }
                            // One fp64 -> fp32 conversion per normal, then the same fp32 split as the generated normals and the tile
                            // kernel's supplied-normals path (hi = rn16(z), lo = rn16(z - hi): z is represented to 2^-22 |z|, the
                            // fp32 rounding adds 2^-24 |z|; the guard band's relative part is 2^-18 or more).  Splitting the double
                            // itself took three fp64 conversions per normal, and the XU pipe they run on bound the kernel (ncu:
                            // 49 % active over the whole launch, profiles/r02_stream_supplied_ncu_summary.txt).
                            float z[16];
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                const double zd = ZTMA ? (col0 + q < kend ? zbuf[q] : 0.0) : ((i < p.M && col0 + q < kend) ? __ldg(zrow + col0 + q) : 0.0);
                                z[q] = (float)zd;
                            }
                            float2 acc2 = make_float2(zsq, 0.0f);
#pragma unroll
                            for (int q = 0; q < 8; ++q) acc2 = __ffma2_rn(make_float2(z[2 * q], z[2 * q + 1]), make_float2(z[2 * q], z[2 * q + 1]), acc2);
                            zsq = acc2.x + acc2.y;
                            split16(z, hv, lv);
                        } else if ((ps > 0 && kb < KCACHE) || (prefetch && kb == KCACHE)) {
                            // generated (and counted into zsq_cached) in the first pass - or, the first K-block after the cached
                            // ones in the last pass, at the beginning of this pass into the cache slot of K-block 0 (below)
                            flush();
                            if (kb == KCACHE) zck = zc;
                            const uint4 h0 = zck[0], h1 = zck[128], l0 = zck[256], l1 = zck[384];
                            hv[0] = h0.x; hv[1] = h0.y; hv[2] = h0.z; hv[3] = h0.w; hv[4] = h1.x; hv[5] = h1.y; hv[6] = h1.z; hv[7] = h1.w;
                            lv[0] = l0.x; lv[1] = l0.y; lv[2] = l0.z; lv[3] = l0.w; lv[4] = l1.x; lv[5] = l1.y; lv[6] = l1.z; lv[7] = l1.w;
                        } else {
                            gen16(kb, kend, path, rng, hv, lv, zsq);
                            if (ps == 0 && kb < KCACHE) {
                                zck[0] = make_uint4(hv[0], hv[1], hv[2], hv[3]); zck[128] = make_uint4(hv[4], hv[5], hv[6], hv[7]);
                                zck[256] = make_uint4(lv[0], lv[1], lv[2], lv[3]); zck[384] = make_uint4(lv[4], lv[5], lv[6], lv[7]);
                                if (kb == KCACHE - 1) zsq_cached = zsq;
                            }
                        }
                    } else {
                        flush();
                    }
                    // stay within ST_ZSLOTS slots of the tensor core
                    if (warp == 0) ST_TRACE(0, 2, kb); else if (warp == 12) ST_TRACE(1, 2, kb);      // normals ready, waiting for the slot
                    if (zu >= (uint32_t)ST_ZSLOTS) mbar_wait<64>(&s_zempty[slot], ((zu / ST_ZSLOTS) - 1u) & 1u);
                    tc_fence_after();
                    if (warp == 0) ST_TRACE(0, 3, kb); else if (warp == 12) ST_TRACE(1, 3, kb);      // slot free
                    // The slot is free, so the MMAs of this pass's K-block kb - 2 are complete; with a lower-triangular transform the
                    // K-blocks after it only touch the rows >= 64 (kb - 1), so the accumulator columns below that are final and are read
                    // now, while the tensor core works on.  The end of the pass is then left with the last few columns: neither do the
                    // generators wait long for the tensor core to drain, nor the tensor core for the read-out.  (The previous pass
                    // is pending until this pass's K-block 0 has been queued: nothing is read early before kb = 2.)
                    if (LOWER && st >= 2 && stream_ascending(ps, KCACHE)) read_columns(ps, min(p.pass[ps].W, 64 * (st - 1) - p.pass[ps].r0), false, i);
                    if (kb >= 0) {
                        tmem_st8(zaddr, hv);
                        tmem_st8(zaddr + 32u, lv);
                    } else if (part == 0) {
                        const uint32_t ones[8] = {0x3C003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};   // 1.0 in K columns 0 and 1
                        tmem_st8(zaddr, ones);
                    }
                    deferred = true; def_slot = slot;
                    if (!MCR_STREAM_DEFER) flush();
                    if (warp == 0) ST_TRACE(0, 4, kb); else if (warp == 12) ST_TRACE(1, 4, kb);      // handed over (or stores issued)
                    // the previous pass is read out once this pass's constant slot and first K-block are queued for the tensor core
                    if (st == 0 && pending) {
                        flush();
                        if (warp == 0) ST_TRACE(0, 5, pend_ps);
                        epilogue(pend_ps, pend_zsq, pend_i, pend_ps == NPASS - 1, true);
                        pending = false;
                        if (warp == 0) ST_TRACE(0, 6, pend_ps);
                    }
                    // Last pass: its first KCACHE K-blocks come from the cache and cost the generators nothing, while their MMAs are
                    // the widest of the pass - the generators would wait for Z slots with their next real work (K-block KCACHE)
                    // still ahead.  That block is generated now, into the cache slot of K-block 0 (read for the last time in this
                    // step, by this very thread), and taken from there when its turn comes.
                    if (prefetch && st == 0) {
                        uint32_t ph[8], pl[8];
                        gen16(KCACHE, kend, path, rng, ph, pl, zsq);
                        zc[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]); zc[128] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                        zc[256] = make_uint4(pl[0], pl[1], pl[2], pl[3]); zc[384] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
                    }
                }
                pending = true; pend_ps = ps; pend_zsq = zsq; pend_i = i;
            }
        }
        flush();
        if (pending) epilogue(pend_ps, pend_zsq, pend_i, true, false);
        if (PORT) {   // warp sums -> four double atomics per warp (the threads that never saw the portfolio column add zeros)
            float a0 = (float)p_hits, a1 = p_sum, a2 = p_sum2, a3 = p_tail;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
            }
            if (lane == 0 && (a0 != 0.0f || a1 != 0.0f || a2 != 0.0f)) {
                atomicAdd(&p.port_acc[0], (double)a0); atomicAdd(&p.port_acc[1], (double)a1);
                atomicAdd(&p.port_acc[2], (double)a2); atomicAdd(&p.port_acc[3], (double)a3);
            }
        }
        if (part == 0) {
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            sum_s = __reduce_add_sync(0xffffffffu, sum_s);
            if (lane == 0) { s_cnt[quarter] = cnt; s_sum[quarter] = sum_s; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (csz > 1) cluster_sync_all();   // nothing of the partner's is still in flight towards this CTA's shared memory
    if (PORT)
        for (int t = tid; t < MCR_PORTFOLIO_BINS; t += ST_THREADS) port_hist_flush(s_hist[t], t, p.hist_lo_d, p.hist_w_d, p.port_hist, p.port_hist_sum);
    if (tid == 0) {
        uint64_t c = 0, s = 0;
        for (int w = 0; w < 4; ++w) { c += s_cnt[w]; s += s_sum[w]; }
        p.block_partials[2 * (uint64_t)blockIdx.x] = my_tiles ? c : 0;
        p.block_partials[2 * (uint64_t)blockIdx.x + 1] = my_tiles ? s : 0;
    }
    if (warp == ST_MMA_WARP) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ---- host side -----------------------------------------------------------------------------------------
static int round16(int v) { return (v + 15) & ~15; }

// passes over row ranges of at most STREAM_WMAX rows: the first pass takes the remainder, so that with a lower-triangular
// transform (pass p needs the normals of columns [0, last row of p]) the re-generated columns are as few as possible
int stream_plan(int N, int port, bool lower, StreamPass* pass) {
    const int rows = round16(N + port);
    const int np = (rows + ST_WMAX - 1) / ST_WMAX;
    if (np > STREAM_MAX_PASSES) return 0;
    int r0 = 0;
    for (int q = 0; q < np; ++q) {
        const int W = q == 0 ? rows - (np - 1) * ST_WMAX : ST_WMAX;
        pass[q].r0 = r0; pass[q].W = W;
        pass[q].kend = lower ? std::min(N, r0 + W) : N;   // normal columns (the portfolio row, last, needs all N)
        pass[q].kbn = (pass[q].kend + tc::KBLOCK - 1) / tc::KBLOCK;
        r0 += W;
    }
    return np;
}

bool stream_supported(int N, int port) { return N >= 1 && round16(N + port) <= ST_WMAX * STREAM_MAX_PASSES && N < MCR_MAX_HIST; }

size_t stream_image_bytes(int N, int port, bool lower) {
    StreamPass ps[STREAM_MAX_PASSES];
    const int np = stream_plan(N, port, lower, ps);
    size_t b = 0;
    for (int q = 0; q < np; ++q) {
        b += (size_t)ps[q].W * 128;
        for (int kb = 0; kb < ps[q].kbn; ++kb) b += 2 * (size_t)(ps[q].W - stream_acol0(ps[q].r0, kb, lower)) * 128;
    }
    return b;
}

double stream_mma_flops(int N, int port, bool lower) {
    StreamPass ps[STREAM_MAX_PASSES];
    const int np = stream_plan(N, port, lower, ps);
    double f = 0;
    for (int q = 0; q < np; ++q) {
        f += 2.0 * 16.0 * ps[q].W;
        for (int kb = 0; kb < ps[q].kbn; ++kb) f += 3.0 * 2.0 * 64.0 * (ps[q].W - stream_acol0(ps[q].r0, kb, lower));
    }
    return f;
}

int stream_normals_per_trial(int N, int port, bool lower) {
    StreamPass ps[STREAM_MAX_PASSES];
    const int np = stream_plan(N, port, lower, ps);
    const int cached = np > 1 ? tc::KBLOCK * std::min(ST_KCACHE_MAX, ps[0].kend / tc::KBLOCK) : 0;
    int s = 0;
    for (int q = 0; q < np; ++q) s += ps[q].kend - (q > 0 ? cached : 0);
    return s;
}

// CTA-pair form (cta_group::2 MMAs, see the kernel): off unless MCR_STREAM_PAIR=1; generated normals without the portfolio row only
bool stream_pair(bool supplied, bool portfolio) {
    static const int want = [] { const char* v = getenv("MCR_STREAM_PAIR"); return v && *v ? atoi(v) : 0; }();
    return want != 0 && !supplied && !portfolio;
}

// element (row r of n, K column kk) of the piece at byte offset `off` of the image.  pair: the image is two rank images of
// total / 2 bytes; rank q holds rows [q n / 2, (q + 1) n / 2) of every piece, at half the piece's offset
static void put_half(uint8_t* image, size_t total, bool pair, size_t off, int n, int r, int kk, __half v) {
    uint8_t* dst;
    if (pair) {
        const int nh = n / 2, q = r / nh;
        dst = image + (size_t)q * (total / 2) + off / 2 + tc::operand_offset(0, r - q * nh, kk);
    } else {
        dst = image + off + tc::operand_offset(0, r, kk);
    }
    memcpy(dst, &v, 2);
}

// image: per pass the constant pieces, then per K-block the hi pieces and the lo pieces, one piece per chunk of <= 256 active rows
// (stream_n1), each (rows) x 64 fp16 in the K-major 128-byte-swizzled operand layout; row i carries s_i A_ij (row_scale:
// tc_row_scale), the constant piece s_i (mu_i + x) as hi (K column 0) and lo (K column 1); padded rows hold zeros and a large
// positive constant
void stream_build(int N, bool lower, const double* A, const double* c, const double* port_row, bool supplied_order, bool pair, uint8_t* image, double* row_scale) {
    StreamPass ps[STREAM_MAX_PASSES];
    const int port = port_row ? 1 : 0;
    const int np = stream_plan(N, port, lower, ps);
    const size_t total = stream_image_bytes(N, port, lower);
    memset(image, 0, total);
    // row i < N: transform row i and the constant mu_i + x; row N (portfolio): w^T A and the constant w.mu
    auto entry = [&](int i, int j) { return i < N ? A[(size_t)i * N + j] : port_row[j]; };
    auto constant = [&](int i) { return i < N ? c[i] : port_row[N]; };
    const int rows = N + port;
    for (int i = 0; i < rows; ++i) {
        double mx = std::fabs(constant(i));
        for (int k = 0; k < N; ++k) mx = std::max(mx, std::fabs(entry(i, k)));
        row_scale[i] = tc_row_scale(mx);
    }
    auto split = [](double v, __half& hi, __half& lo) {
        hi = __float2half_rn((float)v);
        lo = __float2half_rn((float)(v - (double)__half2float(hi)));
    };
    size_t out = 0;
    for (int q = 0; q < np; ++q) {
        const int r0 = ps[q].r0, W = ps[q].W;
        const int kcache = np > 1 ? std::min(ST_KCACHE_MAX, ps[0].kend / tc::KBLOCK) : 0;   // as stream_launch sets it (generated normals)
        for (int st = -1; st < ps[q].kbn; ++st) {
            const int kb = st < 0 ? -1 : stream_kb(q, st, ps[q].kbn, supplied_order ? 0 : kcache);
            const int a0 = kb < 0 ? 0 : stream_acol0(r0, kb, lower), w = W - a0, n1 = stream_n1(w);
            // piece of (half, chunk): halves outer, chunks inner, as the kernel consumes them
            const size_t piece[2][2] = {{out, out + (size_t)n1 * 128}, {out + (size_t)w * 128, out + (size_t)(w + n1) * 128}};
            for (int r = 0; r < w; ++r) {
                const int i = r0 + a0 + r, ch = r >= n1 ? 1 : 0, rl = r - (ch ? n1 : 0), nch = ch ? w - n1 : n1;
                __half hi, lo;
                if (kb < 0) {
                    split(i < rows ? constant(i) * row_scale[i] : TC_PAD_ROW, hi, lo);
                    put_half(image, total, pair, piece[0][ch], nch, rl, 0, hi);
                    put_half(image, total, pair, piece[0][ch], nch, rl, 1, lo);
                    continue;
                }
                if (i >= rows) break;
                for (int kk = 0; kk < tc::KBLOCK; ++kk) {
                    const int j = tc::KBLOCK * kb + kk;
                    if (j >= ps[q].kend) break;
                    split(entry(i, j) * row_scale[i], hi, lo);
                    put_half(image, total, pair, piece[0][ch], nch, rl, kk, hi);
                    put_half(image, total, pair, piece[1][ch], nch, rl, kk, lo);
                }
            }
            out += (size_t)w * 128 * (kb < 0 ? 1 : 2);
        }
    }
}

// one CTA per SM; an even count (CTA pairs share the transform stream) unless there is a single tile
#if MCR_STREAM_TRACE
extern "C" int mcr_stream_trace_dump(unsigned long long* out /* TRACE_ROLES * TRACE_CAP */, unsigned int* counts /* TRACE_ROLES */, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_stream_trace, sizeof(unsigned long long) * TRACE_ROLES * TRACE_CAP);
    cudaMemcpyFromSymbol(counts, g_stream_trace_n, sizeof(unsigned int) * TRACE_ROLES);
    if (reset) { unsigned int z[TRACE_ROLES] = {}; cudaMemcpyToSymbol(g_stream_trace_n, z, sizeof(z)); }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
#endif

int stream_grid(uint64_t M, int num_sms, bool pair) {
    const uint64_t tiles = (M + tc::TILE_M - 1) / tc::TILE_M;
    int g = (int)std::min<uint64_t>((uint64_t)num_sms, tiles ? tiles : 1);
    if (g > 1) g &= ~1;
    if (pair && g < 2) g = 2;   // a CTA pair always; the second CTA of a single-tile run works on a ghost tile
    return g;
}

// tensor map of the caller's normals: fp64, M rows of N, boxes of 16 columns x 128 rows, 128-byte swizzle, zero fill outside
static bool encode_z_map(const double* Z, uint64_t M, int N, CUtensorMap* map) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const encode_fn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        return reinterpret_cast<encode_fn>(fn);
    }();
    if (!encode || (N & 1) || (reinterpret_cast<uintptr_t>(Z) & 15) || M == 0) return false;   // rows must start 16-byte aligned
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 8};
    const cuuint32_t box[2] = {16, 128}, estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(Z), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t stream_launch(const StreamLaunch& l, int num_sms, cudaStream_t s) {
    StreamParams p{};
    p.M = l.M; p.path_offset = l.path_offset; p.key0 = l.key0; p.key1 = l.key1;
    p.N = l.N; p.k = l.k; p.lower = l.lower;
    p.npass = stream_plan(l.N, l.portfolio ? 1 : 0, l.lower != 0, p.pass);
    if (p.npass <= 0) return cudaErrorInvalidValue;
    for (int q = 0; q < p.npass; ++q) {
        p.delta_rel[q] = (float)l.delta_rel[q];
        p.delta_abs[q] = (float)l.delta_abs[q];
        // upper bound of the band over all trials: |z_j| < 7 (Box-Muller on 32-bit uniforms); supplied normals: no such bound
        p.thr_max[q] = l.d_Z ? 3.0e38f : p.delta_rel[q] * 7.0f * sqrtf((float)p.pass[q].kend) * 1.001f + p.delta_abs[q];
    }
    p.image = reinterpret_cast<const uint8_t*>(l.d_image); p.Z = l.d_Z;
    p.S_out = l.S_out; p.flag_list = l.flag_list; p.flag_count = l.flag_count; p.flag_cap = l.flag_cap;
    p.pair_list = l.pair_list; p.pair_count = l.pair_count; p.pair_cap = l.pair_cap; p.block_partials = l.block_partials;
    // K-blocks of normals that lie entirely inside the first pass are kept in shared memory for the later passes; the ring of
    // transform pieces gets the rest of 6 x 32 KB
    p.kcache = (p.npass > 1 && !l.d_Z) ? std::min(ST_KCACHE_MAX, p.pass[0].kend / tc::KBLOCK) : 0;
    p.ring = (ST_RING_MAX - p.kcache) * (l.pair ? 2 : 1);   // (pair: slots of half the size)
    p.image_rank_bytes = stream_image_bytes(l.N, l.portfolio ? 1 : 0, l.lower != 0) / 2;
    alignas(64) CUtensorMap zmap;
    memset(&zmap, 0, sizeof(zmap));
    static const bool no_ztma = [] { const char* v = getenv("MCR_STREAM_NO_ZTMA"); return v && *v && atoi(v) != 0; }();   // A/B runs
    if (l.d_Z && !no_ztma && encode_z_map(l.d_Z, l.M, l.N, &zmap)) {
        p.ztma = 1;
        p.zbox_shift = p.npass > 1 ? 2 : 3;   // one box per part and four transform pieces, or two and two (see ST_ZBOX_MAX)
        p.ring = p.npass > 1 ? 4 : 2;
    }
    const size_t smem = p.ztma ? (size_t)p.ring * ST_PIECE_MAX + ((size_t)ST_ZBOX_BYTES << p.zbox_shift) + 1024 : (size_t)(ST_RING_MAX - p.kcache) * ST_PIECE_MAX + (size_t)p.kcache * ST_ZCACHE_BYTES + 1024 + (l.portfolio ? (size_t)MCR_PORTFOLIO_BINS * 8 : 0);
    p.portfolio = l.portfolio;
    if (l.portfolio) {
        p.port_col = l.N - p.pass[p.npass - 1].r0;
        p.port_thresh = (float)(-l.port_x * l.port_scale); p.port_inv_scale = (float)(1.0 / l.port_scale);
        p.hist_lo = (float)l.hist_lo; p.hist_inv_w = (float)(1.0 / l.hist_width); p.hist_lo_d = l.hist_lo; p.hist_w_d = l.hist_width;
        p.port_acc = l.port_acc; p.port_hist = l.port_hist; p.port_hist_sum = l.port_hist_sum;
    }
    p.keys.set(l.key0, l.key1);
    const int grid = stream_grid(l.M, num_sms, l.pair != 0);
    if (l.pair && (l.d_Z || l.portfolio)) return cudaErrorInvalidValue;
    static const int want_cluster = [] { const char* v = getenv("MCR_STREAM_CLUSTER"); return v && *v ? atoi(v) : 2; }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(l.d_Z ? ST_THREADS_SUPPLIED : ST_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (l.pair || (want_cluster == 2 && grid % 2 == 0)) ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e;
    if (l.d_Z) {
        e = cudaFuncSetAttribute(tc_stream_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaLaunchKernelEx(&cfg, tc_stream_kernel<true, false, false>, p, zmap);
    } else if (l.portfolio) {
        e = cudaFuncSetAttribute(tc_stream_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaLaunchKernelEx(&cfg, tc_stream_kernel<false, true, false>, p, zmap);
    } else if (l.pair) {
        e = cudaFuncSetAttribute(tc_stream_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaLaunchKernelEx(&cfg, tc_stream_kernel<false, false, true>, p, zmap);
    } else {
        e = cudaFuncSetAttribute(tc_stream_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaLaunchKernelEx(&cfg, tc_stream_kernel<false, false, false>, p, zmap);
    }
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mcr
