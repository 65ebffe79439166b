// Counter-based normal generation kept in registers: Philox4x32-10 + Box-Muller.
//
// Replaces the reference GPU model's per-thread cuRAND XORWOW state
// (src/05-GPU/monte_carlos_cuda.cu:30-31,45-47: curand_init with a per-thread subsequence,
// curand_normal_double) and plays the role of gsl_ran_gaussian over MT19937 in the CPU models
// (src/02-C-serial/monte_carlo_serial.c:67-69).  Nothing is stored: the 4 normals of asset
// block b of path p are a pure function of (seed, p, b), which is what makes results
// independent of grid shape and of how the path range is sharded over GPUs.
//
//   key     = (seed_lo, seed_hi)
//   counter = (path_lo, path_hi, b, 0)          -> 4 x u32 -> normals of assets 4b .. 4b+3
//   (u0,u1) -> z0 = r cos t, z1 = r sin t ; (u2,u3) -> z2, z3
//   r = sqrt(-2 ln(u0 * 2^-32 + 2^-33)),  t = 2 pi u1 2^-32       (MUFU lg2/sqrt/sin/cos)
//
// Every float operation is an explicit intrinsic (no contraction freedom), so the main kernel,
// the fix-up kernel and the dump kernel produce bit-identical normals.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mcr {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

// |z| can never exceed this: r <= sqrt(-2 ln 2^-33) = 6.7637; MUFU error margin included.
constexpr float BOX_MULLER_ZMAX = 6.8f;

struct u32x4 { uint32_t x, y, z, w; };

// the round keys of Philox4x32-10: k0[r] = key0 + r * W0, k1[r] = key1 + r * W1 (host-computable)
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
    __host__ __device__ void set(uint32_t key0, uint32_t key1) {
        for (int r = 0; r < 10; ++r) { k0[r] = key0 + (uint32_t)r * PHILOX_W0; k1[r] = key1 + (uint32_t)r * PHILOX_W1; }
    }
};

__device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;   // IMAD.WIDE.U32
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;  // LOP3
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += PHILOX_W0;  // warp-uniform: folded into constants / uniform registers
        k1 += PHILOX_W1;
    }
    return u32x4{c0, c1, c2, c3};
}

__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// two uniforms -> two standard normals
__device__ __forceinline__ void box_muller(uint32_t ua, uint32_t ub, float& za, float& zb) {
    const float u = __fmaf_rn(__uint2float_rn(ua), 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (0, 1]
    const float r = mufu_sqrt(__fmul_rn(-1.3862943611198906f, mufu_lg2(u)));                              // sqrt(-2 ln u)
    const float t = __fmul_rn(__uint2float_rn(ub), 1.4629180792671596e-9f);                               // 2 pi ub 2^-32
    za = __fmul_rn(r, mufu_cos(t));
    zb = __fmul_rn(r, mufu_sin(t));
}

// sqrt(2 ln 2): box_muller's radius is BOX_MULLER_S * sqrt(-lg2 u)
constexpr double BOX_MULLER_S = 1.1774100225154747;
// two uniforms -> two standard normals divided by BOX_MULLER_S, and their common radius r' = sqrt(-lg2 u) >= |za'|, |zb'|
__device__ __forceinline__ void box_muller_raw(uint32_t ua, uint32_t ub, float& za, float& zb, float& r) {
    const float u = __fmaf_rn(__uint2float_rn(ua), 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (0, 1]
    r = mufu_sqrt(-mufu_lg2(u));
    const float t = __fmul_rn(__uint2float_rn(ub), 1.4629180792671596e-9f);
    za = __fmul_rn(r, mufu_cos(t));
    zb = __fmul_rn(r, mufu_sin(t));
}

// the 4 normals of asset block b of one path
__device__ __forceinline__ void normals4(uint32_t path_lo, uint32_t path_hi, uint32_t b,
                                         uint32_t k0, uint32_t k1, float z[4]) {
    const u32x4 u = philox4x32_10(path_lo, path_hi, b, 0u, k0, k1);
    box_muller(u.x, u.y, z[0], z[1]);
    box_muller(u.z, u.w, z[2], z[3]);
}

// Same stream, with the per-path part of Philox hoisted.  With counter (path_lo, path_hi, b, 0):
//   round 1:  p0 = M0 * path_lo  does not depend on b;  p1 = M1 * b  is warp-uniform (uniform datapath);
//   round 2:  p1' = M1 * (hi(p0) ^ k1)  does not depend on b either.
// A rolled loop over b therefore needs 17 vector 32x32->64 multiplies per block instead of 20 (the
// multiply runs only on the heavy FMA pipe, the busiest pipe of the generator).  Bit-identical to
// philox4x32_10 by construction; checked against the CPU Philox in the tests.
struct PhiloxPath {
    uint32_t path_hi, n3, p1p_hi, p1p_lo;
    __device__ __forceinline__ void init(uint32_t path_lo, uint32_t path_hi_, uint32_t k1) {
        path_hi = path_hi_;
        const uint64_t p0 = (uint64_t)PHILOX_M0 * path_lo;
        n3 = (uint32_t)p0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ k1;   // c3 = 0
        const uint64_t p1p = (uint64_t)PHILOX_M1 * n2;
        p1p_hi = (uint32_t)(p1p >> 32);
        p1p_lo = (uint32_t)p1p;
    }
    __device__ __forceinline__ u32x4 block(uint32_t b, uint32_t k0, uint32_t k1) const {
        // round 1 (key k0, k1)
        const uint64_t p1 = (uint64_t)PHILOX_M1 * b;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ path_hi ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        // round 2 (key k0 + W0, k1 + W1): c = (n0, n1, n2, n3)
        k0 += PHILOX_W0; k1 += PHILOX_W1;
        const uint64_t q0 = (uint64_t)PHILOX_M0 * n0;
        uint32_t c0 = p1p_hi ^ n1 ^ k0;
        uint32_t c1 = p1p_lo;
        uint32_t c2 = (uint32_t)(q0 >> 32) ^ n3 ^ k1;
        uint32_t c3 = (uint32_t)q0;
#pragma unroll
        for (int r = 2; r < 10; r++) {
            k0 += PHILOX_W0; k1 += PHILOX_W1;
            const uint64_t a0 = (uint64_t)PHILOX_M0 * c0;
            const uint64_t a1 = (uint64_t)PHILOX_M1 * c2;
            const uint32_t m0 = (uint32_t)(a1 >> 32) ^ c1 ^ k0;
            const uint32_t m2 = (uint32_t)(a0 >> 32) ^ c3 ^ k1;
            c1 = (uint32_t)a1;
            c3 = (uint32_t)a0;
            c0 = m0;
            c2 = m2;
        }
        return u32x4{c0, c1, c2, c3};
    }
    // NB blocks b, b+1, ... in lock step: the rounds of the blocks are interleaved in program order, so
    // the instruction stream carries 2*NB independent multiply chains instead of 2 (ptxas otherwise
    // schedules the blocks one after the other and every round waits out the multiply latency)
    template <int NB>
    __device__ __forceinline__ void normals_multi(uint32_t b, uint32_t k0, uint32_t k1, float* z) const {
        uint32_t c0[NB], c1[NB], c2[NB], c3[NB];
        uint32_t ka = k0 + PHILOX_W0, kb = k1 + PHILOX_W1;   // round-2 keys
#pragma unroll
        for (int t = 0; t < NB; ++t) {
            const uint64_t p1 = (uint64_t)PHILOX_M1 * (b + t);
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ path_hi ^ k0;
            const uint32_t n1 = (uint32_t)p1;
            const uint64_t q0 = (uint64_t)PHILOX_M0 * n0;
            c0[t] = p1p_hi ^ n1 ^ ka;
            c1[t] = p1p_lo;
            c2[t] = (uint32_t)(q0 >> 32) ^ n3 ^ kb;
            c3[t] = (uint32_t)q0;
        }
#pragma unroll
        for (int r = 2; r < 10; r++) {
            ka += PHILOX_W0; kb += PHILOX_W1;
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                const uint64_t a0 = (uint64_t)PHILOX_M0 * c0[t];
                const uint64_t a1 = (uint64_t)PHILOX_M1 * c2[t];
                const uint32_t m0 = (uint32_t)(a1 >> 32) ^ c1[t] ^ ka;
                const uint32_t m2 = (uint32_t)(a0 >> 32) ^ c3[t] ^ kb;
                c1[t] = (uint32_t)a1;
                c3[t] = (uint32_t)a0;
                c0[t] = m0;
                c2[t] = m2;
            }
        }
#pragma unroll
        for (int t = 0; t < NB; ++t) {
            box_muller(c0[t], c1[t], z[4 * t], z[4 * t + 1]);
            box_muller(c2[t], c3[t], z[4 * t + 2], z[4 * t + 3]);
        }
    }
    __device__ __forceinline__ void normals(uint32_t b, uint32_t k0, uint32_t k1, float z[4]) const {
        const u32x4 u = block(b, k0, k1);
        box_muller(u.x, u.y, z[0], z[1]);
        box_muller(u.z, u.w, z[2], z[3]);
    }
    // ---- the same functions with the ten round keys precomputed (PhiloxKeys lives in the kernel parameters, so
    // every key is a constant-bank operand of its LOP3 and the key schedule costs no instruction) ----
    __device__ __forceinline__ u32x4 block(uint32_t b, const PhiloxKeys& K) const {
        const uint64_t p1 = (uint64_t)PHILOX_M1 * b;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ path_hi ^ K.k0[0];
        const uint32_t n1 = (uint32_t)p1;
        const uint64_t q0 = (uint64_t)PHILOX_M0 * n0;
        uint32_t c0 = p1p_hi ^ n1 ^ K.k0[1];
        uint32_t c1 = p1p_lo;
        uint32_t c2 = (uint32_t)(q0 >> 32) ^ n3 ^ K.k1[1];
        uint32_t c3 = (uint32_t)q0;
#pragma unroll
        for (int r = 2; r < 10; r++) {
            const uint64_t a0 = (uint64_t)PHILOX_M0 * c0;
            const uint64_t a1 = (uint64_t)PHILOX_M1 * c2;
            const uint32_t m0 = (uint32_t)(a1 >> 32) ^ c1 ^ K.k0[r];
            const uint32_t m2 = (uint32_t)(a0 >> 32) ^ c3 ^ K.k1[r];
            c1 = (uint32_t)a1;
            c3 = (uint32_t)a0;
            c0 = m0;
            c2 = m2;
        }
        return u32x4{c0, c1, c2, c3};
    }
    template <int NB>
    __device__ __forceinline__ void normals_multi(uint32_t b, const PhiloxKeys& K, float* z) const {
        normals_multi_between<NB>(b, K, z, [] {});
    }
    // the same with `between()` run after the Philox rounds and before the Box-Muller transforms (a place to retire an
    // asynchronous operation issued before the call: ~50 instructions per block have gone by)
    template <int NB, class F>
    __device__ __forceinline__ void normals_multi_between(uint32_t b, const PhiloxKeys& K, float* z, F between) const {
        uint32_t c0[NB], c1[NB], c2[NB], c3[NB];
#pragma unroll
        for (int t = 0; t < NB; ++t) {
            const uint64_t p1 = (uint64_t)PHILOX_M1 * (b + t);
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ path_hi ^ K.k0[0];
            const uint32_t n1 = (uint32_t)p1;
            const uint64_t q0 = (uint64_t)PHILOX_M0 * n0;
            c0[t] = p1p_hi ^ n1 ^ K.k0[1];
            c1[t] = p1p_lo;
            c2[t] = (uint32_t)(q0 >> 32) ^ n3 ^ K.k1[1];
            c3[t] = (uint32_t)q0;
        }
#pragma unroll
        for (int r = 2; r < 10; r++) {
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                const uint64_t a0 = (uint64_t)PHILOX_M0 * c0[t];
                const uint64_t a1 = (uint64_t)PHILOX_M1 * c2[t];
                const uint32_t m0 = (uint32_t)(a1 >> 32) ^ c1[t] ^ K.k0[r];
                const uint32_t m2 = (uint32_t)(a0 >> 32) ^ c3[t] ^ K.k1[r];
                c1[t] = (uint32_t)a1;
                c3[t] = (uint32_t)a0;
                c0[t] = m0;
                c2[t] = m2;
            }
        }
        between();
#pragma unroll
        for (int t = 0; t < NB; ++t) {
            box_muller(c0[t], c1[t], z[4 * t], z[4 * t + 1]);
            box_muller(c2[t], c3[t], z[4 * t + 2], z[4 * t + 3]);
        }
    }
    __device__ __forceinline__ void normals(uint32_t b, const PhiloxKeys& K, float z[4]) const {
        const u32x4 u = block(b, K);
        box_muller(u.x, u.y, z[0], z[1]);
        box_muller(u.z, u.w, z[2], z[3]);
    }
    // only the first Box-Muller pair of block b (the last asset block of a path when N % 4 is 1 or 2)
    __device__ __forceinline__ void normals_pair(uint32_t b, const PhiloxKeys& K, float z[2]) const {
        const u32x4 u = block(b, K);
        box_muller(u.x, u.y, z[0], z[1]);
    }
    // "Raw" flavour for the tensor kernel: z' = z / sqrt(2 ln 2) = sqrt(-lg2 u) (cos t, sin t); the caller folds the factor
    // into the matrix that multiplies z' (one FMUL per pair less) and bounds |z'| by the largest radius seen, rmax (one
    // FMNMX3 per Philox block instead of two).  LAST_NORMALS = how many normals of the last block are needed (2 or 4).
    template <int NB, int LAST_NORMALS = 4>
    __device__ __forceinline__ void normals_multi_raw(uint32_t b, const PhiloxKeys& K, float* z, float& rmax) const {
        uint32_t c0[NB], c1[NB], c2[NB], c3[NB];
#pragma unroll
        for (int t = 0; t < NB; ++t) {
            const uint64_t p1 = (uint64_t)PHILOX_M1 * (b + t);
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ path_hi ^ K.k0[0];
            const uint32_t n1 = (uint32_t)p1;
            const uint64_t q0 = (uint64_t)PHILOX_M0 * n0;
            c0[t] = p1p_hi ^ n1 ^ K.k0[1];
            c1[t] = p1p_lo;
            c2[t] = (uint32_t)(q0 >> 32) ^ n3 ^ K.k1[1];
            c3[t] = (uint32_t)q0;
        }
#pragma unroll
        for (int r = 2; r < 10; r++) {
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                const uint64_t a0 = (uint64_t)PHILOX_M0 * c0[t];
                const uint64_t a1 = (uint64_t)PHILOX_M1 * c2[t];
                const uint32_t m0 = (uint32_t)(a1 >> 32) ^ c1[t] ^ K.k0[r];
                const uint32_t m2 = (uint32_t)(a0 >> 32) ^ c3[t] ^ K.k1[r];
                c1[t] = (uint32_t)a1;
                c3[t] = (uint32_t)a0;
                c0[t] = m0;
                c2[t] = m2;
            }
        }
#pragma unroll
        for (int t = 0; t < NB; ++t) {
            float ra, rb = 0.0f;
            box_muller_raw(c0[t], c1[t], z[4 * t], z[4 * t + 1], ra);
            if (t < NB - 1 || LAST_NORMALS > 2) box_muller_raw(c2[t], c3[t], z[4 * t + 2], z[4 * t + 3], rb);
            rmax = fmaxf(rmax, fmaxf(ra, rb));
        }
    }
};

}  // namespace mcr
