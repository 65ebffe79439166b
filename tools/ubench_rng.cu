// Micro-benchmark: Philox4x32-10 + Box-Muller normals per second vs resident warps and blocks in flight.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gpu-montecarlo-risk_b200/csrc -o tools/ubench_rng tools/ubench_rng.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "mcr_rng.cuh"
using namespace mcr;

template <int NBLK>   // Philox blocks interleaved per thread
__global__ void __launch_bounds__(256) k(uint64_t paths, int nb, uint32_t k0, uint32_t k1, float* out) {
    float acc = 0.f;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < paths; i += (uint64_t)gridDim.x * blockDim.x) {
        PhiloxPath rng;
        rng.init((uint32_t)i, (uint32_t)(i >> 32), k1);
        for (int b = 0; b + NBLK <= nb; b += NBLK) {
            float z[4 * NBLK];
            rng.normals_multi<NBLK>((uint32_t)b, k0, k1, z);
#pragma unroll
            for (int q = 0; q < 4 * NBLK; ++q) acc += z[q];
        }
    }
    if (acc == 1234.5f) out[0] = acc;
}

template <int NBLK>
void run(int warps_per_sm) {
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int nb = 12;                       // 48 normals per path
    const uint64_t paths = 40000000;
    const int grid = 148 * warps_per_sm / 8; // 256-thread blocks
    k<NBLK><<<grid, 256>>>(paths / 10, nb, 67, 0, out);
    cudaEventRecord(e0);
    k<NBLK><<<grid, 256>>>(paths, nb, 67, 0, out);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double normals = (double)paths * nb * 4;
    printf("blocks interleaved %d, %2d warps/SM: %8.3f ms  %.3e normals/s  (%.1f SMSP-cycles per warp-normal at 1965 MHz)\n",
           NBLK, warps_per_sm, ms, normals / (ms * 1e-3), ms * 1e-3 * 1.965e9 * 592 / (normals / 32));
}

int main() {
    for (int w : {8, 16, 24, 32, 48, 64}) run<1>(w);
    for (int w : {8, 16, 24, 32, 48, 64}) run<2>(w);
    for (int w : {8, 16, 32, 64}) run<4>(w);
    return 0;
}
