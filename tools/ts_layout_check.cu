// Correctness check of the TMEM-resident A operand (".ts" form of tcgen05.mma.kind::f16) as the streamed N > 128 kernel uses it:
//   * A (M = 128 rows = trials, K = 64 = four K-steps) is written by the threads with tcgen05.st.32x32b: lane = row, TMEM column
//     c of the operand = the fp16 pair (k = 2c, k = 2c + 1), low half first;
//   * B (N rows x 64, K-major, 128-byte swizzle) comes from shared memory through a descriptor whose start address is moved to an
//     arbitrary 8-row group (the kernel multiplies the rows >= the K-block's first column only);
//   * D (fp32, 128 lanes x N columns) lands at an arbitrary TMEM column offset.
// Prints the largest |D - A B^T| against a host fp64 product (exact products of small integers: must be 0).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gpu-montecarlo-risk_b200/csrc -o tools/ts_layout_check tools/ts_layout_check.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "mcr_tc_common.cuh"
using namespace mcr;
using namespace mcr::tc;

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

// NB = rows of B in shared memory, N = UMMA N (rows row0 .. row0 + N - 1 are multiplied), dcol = accumulator column offset
__global__ void __launch_bounds__(128, 1) ts_check(const __half* A, const uint8_t* B_image, int NB, int N, int row0, int dcol, int acol, float* D) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < NB * 128 / 16; i += 128) ((uint4*)smem)[i] = ((const uint4*)B_image)[i];
    if (tid == 0) mbar_init(&s_bar, 1);
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    // A: row = tid, 64 fp16 -> 32 packed columns at TMEM column acol, four x8 stores (one per K-step)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t v[8];
        for (int c = 0; c < 8; ++c) {
            const __half lo = A[tid * 64 + 16 * ks + 2 * c], hi = A[tid * 64 + 16 * ks + 2 * c + 1];
            v[c] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
        }
        tmem_st8(lane_addr + (uint32_t)(acol + 8 * ks), v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        if (elect_one()) {
            tc_fence_after();
            const uint64_t b = make_desc(smem_u32(smem + (row0 / 8) * 1024));
            const uint32_t idesc = make_idesc(N);
            for (int ks = 0; ks < 4; ++ks) umma_f16_ts(tmem_base + dcol, tmem_base + acol + 8 * ks, b + 2 * ks, idesc, ks ? 1u : 0u);
            umma_commit(&s_bar);
        }
        __syncwarp();
    }
    mbar_wait(&s_bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + (uint32_t)(dcol + c0), v);
        tmem_ld_wait();
        for (int c = 0; c < 32 && c0 + c < N; ++c) D[tid * N + c0 + c] = __uint_as_float(v[c]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static int run(int NB, int N, int row0, int dcol, int acol) {
    std::vector<__half> A(128 * 64);
    std::vector<float> Af(128 * 64), Bf((size_t)NB * 64);
    std::vector<uint8_t> img((size_t)NB * 128, 0);
    srand(1234 + N + row0);
    for (int i = 0; i < 128 * 64; ++i) { Af[i] = (float)(rand() % 17 - 8); A[i] = __float2half(Af[i]); }
    for (int r = 0; r < NB; ++r)
        for (int k = 0; k < 64; ++k) {
            Bf[(size_t)r * 64 + k] = (float)(rand() % 13 - 6);
            const __half h = __float2half(Bf[(size_t)r * 64 + k]);
            memcpy(img.data() + operand_offset(NB, r, k), &h, 2);
        }
    __half* dA; uint8_t* dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, img.size()); cudaMalloc(&dD, sizeof(float) * 128 * N);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, img.data(), img.size(), cudaMemcpyHostToDevice);
    const size_t smem = (size_t)NB * 128 + 1024;
    cudaFuncSetAttribute(ts_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ts_check<<<1, 128, smem>>>(dA, dB, NB, N, row0, dcol, acol, dD);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D((size_t)128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < N; ++j) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += (double)Af[i * 64 + k] * Bf[(size_t)(row0 + j) * 64 + k];
            const double d = fabs(ref - D[(size_t)i * N + j]);
            if (d > worst) worst = d;
        }
    printf("B rows %3d  N=%3d  first row %3d  D column %3d  A column %3d : max |D - A B^T| = %g  %s %s\n", NB, N, row0, dcol, acol, worst,
           worst == 0 ? "OK" : "MISMATCH", e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return worst == 0 && e == cudaSuccess ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run(64, 64, 0, 0, 256);
    bad += run(256, 256, 0, 128, 0);
    bad += run(384, 192, 192, 128 + 192, 64);
    // (the read-back uses 32-column tcgen05.ld: keep D column + N rounded up to 32 inside the 512 columns)
    bad += run(384, 176, 208, 128 + 192, 0);
    bad += run(384, 224, 160, 160, 64);
    bad += run(384, 16, 368, 480, 64);
    printf(bad ? "FAILED\n" : "all layouts as assumed\n");
    return bad;
}
