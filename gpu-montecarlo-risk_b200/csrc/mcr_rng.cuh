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

// the 4 normals of asset block b of one path
__device__ __forceinline__ void normals4(uint32_t path_lo, uint32_t path_hi, uint32_t b,
                                         uint32_t k0, uint32_t k1, float z[4]) {
    const u32x4 u = philox4x32_10(path_lo, path_hi, b, 0u, k0, k1);
    box_muller(u.x, u.y, z[0], z[1]);
    box_muller(u.z, u.w, z[2], z[3]);
}

}  // namespace mcr
