// tcgen05 / TMEM / mbarrier wrappers and the fp16 hi+lo operand writer shared by the tensor-core kernels
// (mcr_tc.cu: N <= 128 and the row-slab kernel; mcr_wide.cu: the single-pass kernel for N > 128).
#pragma once
#include <cstdint>
#include <cmath>
#include <cuda_fp16.h>

namespace mcr {

namespace tc {

constexpr int TILE_M = 128;        // trials per tile = UMMA M = TMEM lanes
constexpr int KBLOCK = 64;         // fp16 elements per 128-byte swizzle row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// SLEEP_NS > 0: back off that long between failed polls, so that a waiting warp leaves its issue slots to the others
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (SLEEP_NS > 0) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE;\n"
            "WAIT_LOOP:\n"
            "nanosleep.u32 %2;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE;\n"
            "bra WAIT_LOOP;\n"
            "DONE:\n"
            "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)SLEEP_NS) : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_LOOP:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE;\n"
            "bra WAIT_LOOP;\n"
            "DONE:\n"
            "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

// ---- tcgen05 wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 inputs, fp32 accumulate, one CTA
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// one thread of the (converged) warp; a region guarded by this predicate is known to the compiler to run in a single thread,
// so it issues UTCHMMA / UTCBAR straight from uniform registers (guarded by lane == 0 it wraps each one in an ELECT retry loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0u;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"): 8-row groups of
// 128-byte rows, 1024 bytes apart (SBO); the 16-byte chunk index is XORed with (row % 8).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                       // descriptor version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;                       // layout type: SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=n
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4)        // c_format = F32; a_format = b_format = 0 (F16)
           | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// byte offset of element (row r, column k) inside an operand image made of K-blocks of [rows x 64] fp16
__host__ __device__ inline uint32_t operand_offset(int rows, int r, int k) {
    const int kb = k / KBLOCK, kk = k % KBLOCK;
    return (uint32_t)kb * (uint32_t)rows * 128u + (uint32_t)(r / 8) * 1024u + (uint32_t)(r % 8) * 128u +
           (uint32_t)(((kk / 8) ^ (r % 8)) * 16) + (uint32_t)(kk % 8) * 2u;
}

}  // namespace tc

// Scale of one transform row in the fp16 hi + lo operand images: a power of two (so it is exact and signs, the guard band
// and the portfolio row simply carry it) that puts the row's largest |entry| - its transform entries times the generator's
// z factor, and its constant mu_i + x - into [2^12, 2^13).  A fixed scale (round 1: 2^10) overflows fp16 for returns
// expressed in basis points and pushes the lo parts of small-volatility rows into fp16 subnormals, whose 2^-25 absolute
// rounding error the relative guard band does not cover; per row, every row is as well conditioned as fp16 allows whatever
// the units of Sigma.  Accumulator i then holds s_i * (R_i + x).
inline double tc_row_scale(double row_maxabs) {
    if (!(row_maxabs > 0.0)) return 1.0;
    int e = 0;
    (void)frexp(row_maxabs, &e);          // row_maxabs = m * 2^e, m in [0.5, 1)
    return ldexp(1.0, 13 - e);
}
constexpr double TC_PAD_ROW = 4096.0;     // scaled constant of a padded row: far on the "no crash" side

// ---- portfolio histogram (north-star extension A7), shared by the tile kernel and the streamed kernel -------------------------
// A CTA keeps one 64-bit word per bin in shared memory: count << 44 | sum of q, q = the trial's position inside the histogram
// range in units of 2^-24 of the range (<= 2^24, so 2^20 trials fit below the count field: a CTA sees < 2^20 trials per launch).
// Integer adds only: the per-bin sums are exact in that quantisation (one step = 1e-6 sigma_p) and do not depend on the order in
// which trials arrive, unlike the fp32 atomics of round 1.
constexpr int PORT_HIST_BINS = 2048;                       // = MCR_PORTFOLIO_BINS
constexpr float PORT_HIST_Q_PER_BIN = 8192.0f;             // 2^24 / 2048
__device__ __forceinline__ unsigned long long port_hist_word(float P, float hist_lo, float hist_inv_w, int& bin) {
    float u = (P - hist_lo) * hist_inv_w;                  // position in bins
    u = fminf(fmaxf(u, 0.0f), (float)PORT_HIST_BINS);
    bin = min((int)u, PORT_HIST_BINS - 1);
    return (1ull << 44) + (unsigned long long)__float2uint_rn(u * PORT_HIST_Q_PER_BIN);
}
__device__ __forceinline__ void port_hist_flush(unsigned long long word, int bin, double hist_lo, double hist_w, unsigned long long* hist,
                                                double* hist_sum) {
    const unsigned long long c = word >> 44;
    if (c == 0ull) return;
    const double q = (double)(word & ((1ull << 44) - 1ull));
    atomicAdd(&hist[bin], c);
    atomicAdd(&hist_sum[bin], (double)c * hist_lo + q * (hist_w / (double)PORT_HIST_Q_PER_BIN));   // sum of P = sum (lo + q w / 8192)
}

__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

#ifndef MCR_TC_ABLATE
#define MCR_TC_ABLATE 0
#endif

// z - float(h) with h an fp16 value: fma.rn.f32.f16 (h * -1.0h + z), one FHFMA instruction
__device__ __forceinline__ float sub_half(float z, unsigned short h) {
    float d;
    const unsigned short minus_one = 0xBC00;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(h), "h"(minus_one), "f"(z));
    return d;
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// split 8 normals into fp16 hi + lo and store them as one 16-byte chunk (chunk index c8) of the thread's
// row in the K-major, 128-byte-swizzled operand images.  zhi_row / zlo_row: shared-space address of the
// thread's row in K-block 0; r7s = (row % 8) << 4 (the swizzle XOR term).
template <bool TRACK_ZMAX = true>
__device__ __forceinline__ void store_chunk(const float (&z)[8], int c8, uint32_t zhi_row, uint32_t zlo_row, uint32_t r7s, float& zmax) {
    using namespace tc;
#if MCR_TC_ABLATE == 1   // tuning only: RNG without split / store
#pragma unroll
    for (int q = 0; q < 8; ++q) zmax = fmaxf(zmax, fabsf(z[q]));
    if (zmax == 123.0f) sts128(zhi_row, 0, 0, 0, 0);
    return;
#endif
    uint32_t hv[4], lv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (TRACK_ZMAX) zmax = fmaxf(zmax, fmaxf(fabsf(z[2 * q]), fabsf(z[2 * q + 1])));
        const __half2 h = __floats2half2_rn(z[2 * q], z[2 * q + 1]);
        hv[q] = *reinterpret_cast<const uint32_t*>(&h);
        // lo = z - float(hi) in one mixed-precision FMA each (FHFMA: f16 * f16 + f32, exact here)
        const __half2 l = __floats2half2_rn(sub_half(z[2 * q], (unsigned short)(hv[q] & 0xffffu)),
                                            sub_half(z[2 * q + 1], (unsigned short)(hv[q] >> 16)));
        lv[q] = *reinterpret_cast<const uint32_t*>(&l);
    }
    const uint32_t off = (((uint32_t)c8 & 7u) << 4) ^ r7s;   // c8: chunk index inside the K-block
    sts128(zhi_row + off, hv[0], hv[1], hv[2], hv[3]);
    sts128(zlo_row + off, lv[0], lv[1], lv[2], lv[3]);
}

// an all-zero 16-byte chunk (columns beyond N) in both images
__device__ __forceinline__ void store_zero_chunk(int c8, uint32_t zhi_row, uint32_t zlo_row, uint32_t r7s) {
    const uint32_t off = (((uint32_t)c8 & 7u) << 4) ^ r7s;
    sts128(zhi_row + off, 0u, 0u, 0u, 0u);
    sts128(zlo_row + off, 0u, 0u, 0u, 0u);
}

}  // namespace mcr
