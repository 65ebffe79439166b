// Micro-benchmark: issue rate of the Blackwell packed FMA (FFMA2) by operand form.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_ffma2 ubench_ffma2.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__constant__ float2 cK[64];

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t pk(float a, float b) { return (uint64_t)__float_as_uint(a) | ((uint64_t)__float_as_uint(b) << 32); }

constexpr int NACC = 24;

// MODE 0: reg pair * reg pair + acc        1: UR pair * broadcast scalar reg + acc
//      2: reg pair * broadcast scalar + acc 3: UR pair * reg pair + acc
//      4: scalar FFMA, UR operand           5: scalar FFMA, all regs
template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, float* out, float zin) {
    uint64_t acc[NACC];
    float sacc[NACC];
    const float z0 = zin + 1e-6f * threadIdx.x;
#pragma unroll
    for (int q = 0; q < NACC; ++q) { acc[q] = pk(1.0f + q, 2.0f + q); sacc[q] = 1.0f + q; }
    uint64_t areg[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) areg[q] = pk(0.999f + 1e-5f * (threadIdx.x + q), 0.998f + 1e-5f * q);
    const uint64_t* cK64 = reinterpret_cast<const uint64_t*>(cK);
    const float* cKf = reinterpret_cast<const float*>(cK);
    float z = z0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const uint64_t zz = pk(z, z);
        const uint64_t zp = pk(z, z * 1.0001f);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < NACC; ++q) {
                if (MODE == 0) acc[q] = ffma2(areg[q % 8], zp, acc[q]);
                if (MODE == 1) acc[q] = ffma2(cK64[(q + r) % 32], zz, acc[q]);
                if (MODE == 2) acc[q] = ffma2(areg[q % 8], zz, acc[q]);
                if (MODE == 3) acc[q] = ffma2(cK64[(q + r) % 32], zp, acc[q]);
                if (MODE == 4) sacc[q] = fmaf(cKf[(q + r) % 64], z, sacc[q]);
                if (MODE == 5) sacc[q] = fmaf(__uint_as_float((uint32_t)areg[q % 8]), z, sacc[q]);
            }
        z = z * 0.9999f;
    }
    float s = 0;
#pragma unroll
    for (int q = 0; q < NACC; ++q) s += __uint_as_float((uint32_t)acc[q]) + __uint_as_float((uint32_t)(acc[q] >> 32)) + sacc[q];
    if (s == 1234.5f) out[0] = s;
}

template <int MODE>
void run(const char* name, int lanes_per_inst) {
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, iters = 20000;
    k<MODE><<<grid, 256>>>(100, out, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(iters, out, 0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)grid * 256 / 32 * iters * 4.0 * NACC;   // warp instructions
    const double fma = inst * 32 * lanes_per_inst;
    const double per_smsp_clk = inst / (148.0 * 4) / (ms * 1e-3 * 1.965e9);
    printf("%-44s %8.3f ms  %7.2f TFLOP/s  %.3f warp-inst/clk/SMSP (at 1965 MHz)\n", name, ms, 2 * fma / (ms * 1e-3) / 1e12, per_smsp_clk);
}

int main() {
    float2 h[64];
    for (int i = 0; i < 64; ++i) h[i] = make_float2(0.999f + 1e-5f * i, 0.998f + 1e-5f * i);
    cudaMemcpyToSymbol(cK, h, sizeof(h));
    run<0>("FFMA2 reg.pair * reg.pair + acc", 2);
    run<1>("FFMA2 UR.pair * reg.F32(bcast) + acc", 2);
    run<2>("FFMA2 reg.pair * reg.F32(bcast) + acc", 2);
    run<3>("FFMA2 UR.pair * reg.pair + acc", 2);
    run<4>("FFMA  UR * reg + acc", 1);
    run<5>("FFMA  reg * reg + acc", 1);
    return 0;
}
