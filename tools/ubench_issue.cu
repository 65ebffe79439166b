// Micro-benchmark: does the packed FMA (FFMA2) free issue slots for other pipes on B200?
// Each kernel interleaves F packed FMAs with X instructions of another kind per group and reports
// cycles per group per SM sub-partition (at the nominal 1965 MHz).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__constant__ float2 cK[4096];

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t pk(float a, float b) { return (uint64_t)__float_as_uint(a) | ((uint64_t)__float_as_uint(b) << 32); }
__device__ __forceinline__ void ldc128(uint64_t addr, uint64_t& a, uint64_t& b) {
    asm volatile("ld.const.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(addr));
}

constexpr int NACC = 24;
// KIND 0: FFMA2 only; 1: + LOP3 (alu); 2: + LDCU.128 feeding the FFMA2s; 3: + MUFU; 4: + IMAD.WIDE; 5: scalar FFMA + LOP3
template <int KIND, int F, int X>
__global__ void __launch_bounds__(128) k(int iters, float* out, float zin, uint32_t seed) {
    uint64_t acc[NACC];
    float sacc[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) { acc[q] = pk(1.0f + q, 2.0f + q); sacc[q] = 1.0f + q; }
    const float z = zin + 1e-6f * threadIdx.x;
    const uint64_t zz = pk(z, z);
    uint64_t areg = pk(0.999f, 0.998f);
    uint32_t x0 = seed + threadIdx.x, x1 = seed * 3 + 1, x2 = seed * 5 + 2, x3 = seed * 7 + 3;
    float m0 = 1.5f + z, m1 = 2.5f + z;
    uint64_t w0 = x0, w1 = x1;
    const uint64_t cbase = (uint64_t)__cvta_generic_to_constant(cK);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < 16; ++g) {
            uint64_t a0 = areg, a1 = areg;
            if (KIND == 2) ldc128(cbase + 16 * ((g * 7 + 3) % 256), a0, a1);
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const int q = (g * F + f) % NACC;
                if (KIND == 5) { sacc[q] = fmaf(sacc[(q + 1) % NACC] * 0.f + 0.999f, z, sacc[q]); }
                else acc[q] = ffma2((f & 1) ? a1 : a0, zz, acc[q]);
            }
#pragma unroll
            for (int t = 0; t < X; ++t) {
                if (KIND == 1 || KIND == 5) { x0 = (x0 ^ x1) & (x2 | 0x55u); x1 = (x1 ^ x2) | (x3 & 0x33u); x2 ^= x3 + 0; x3 = (x3 & x0) ^ 0x9e3779b9u; }
                if (KIND == 3) { m0 = __sinf(m0); m1 = __cosf(m1); }
                if (KIND == 4) { w0 = (uint64_t)(uint32_t)w0 * 0xD2511F53u + (w1 >> 32); w1 = (uint64_t)(uint32_t)w1 * 0xCD9E8D57u + (w0 >> 32); }
            }
        }
    }
    float s = (float)(x0 ^ x1 ^ x2 ^ x3) + m0 + m1 + (float)(w0 ^ w1);
#pragma unroll
    for (int q = 0; q < NACC; ++q) s += __uint_as_float((uint32_t)acc[q]) + __uint_as_float((uint32_t)(acc[q] >> 32)) + sacc[q];
    if (s == 1234.5f) out[0] = s;
}

template <int KIND, int F, int X>
void run(const char* name) {
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 4, iters = 20000;   // 16 warps / SM = 4 per sub-partition
    k<KIND, F, X><<<grid, 128>>>(100, out, 0.5f, 1u);
    cudaEventRecord(e0);
    k<KIND, F, X><<<grid, 128>>>(iters, out, 0.5f, 1u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double groups_per_smsp = 4.0 /*warps*/ * iters * 16.0;
    printf("%-52s %8.3f ms   %.2f cycles / group / SMSP\n", name, ms, ms * 1e-3 * 1.965e9 / groups_per_smsp);
}

int main() {
    static float2 h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = make_float2(0.999f + 1e-6f * i, 0.998f + 1e-6f * i);
    cudaMemcpyToSymbol(cK, h, sizeof(h));
    run<0, 2, 0>("2 FFMA2");
    run<0, 4, 0>("4 FFMA2");
    run<1, 2, 1>("2 FFMA2 + 4 LOP3");
    run<1, 2, 2>("2 FFMA2 + 8 LOP3");
    run<1, 4, 1>("4 FFMA2 + 4 LOP3");
    run<2, 2, 0>("2 FFMA2 + 1 LDCU.128 (feeds them)");
    run<2, 4, 0>("4 FFMA2 + 1 LDCU.128 (feeds them)");
    run<2, 8, 0>("8 FFMA2 + 1 LDCU.128 (feeds them)");
    run<3, 2, 1>("2 FFMA2 + 2 MUFU");
    run<3, 4, 1>("4 FFMA2 + 2 MUFU");
    run<3, 8, 1>("8 FFMA2 + 2 MUFU");
    run<4, 2, 1>("2 FFMA2 + 2 IMAD.WIDE");
    run<4, 4, 1>("4 FFMA2 + 2 IMAD.WIDE");
    run<5, 4, 1>("4 FFMA + 4 LOP3");
    run<5, 4, 0>("4 FFMA");
    return 0;
}
