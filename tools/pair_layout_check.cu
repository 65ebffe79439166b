// Correctness check of the CTA-pair form of the streamed kernel's MMA (tcgen05.mma.cta_group::2, A in TMEM, B in shared memory):
//   * a cluster of two CTAs; each CTA writes ITS 128 rows of A (64 fp16 = four K-steps) into its own TMEM with tcgen05.st;
//   * each CTA holds HALF of the B rows (N / 2 rows x 64, K-major, 128-byte swizzle) at the same shared-memory offset;
//   * the leader CTA (rank 0) issues the MMAs with M = 256 and N = n and commits with .multicast::cluster to a barrier in both CTAs;
//   * every CTA reads its 128 x N accumulator back and compares with A_cta * B^T, B = [rows of CTA 0 ; rows of CTA 1].
// Also checks the hand-shake the kernel uses: the peer CTA's threads signal "my A tile is written" with a remote
// mbarrier.arrive.release.cluster on the leader's barrier.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gpu-montecarlo-risk_b200/csrc -o tools/pair_layout_check tools/pair_layout_check.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "mcr_tc_common.cuh"
using namespace mcr;
using namespace mcr::tc;

__device__ __forceinline__ void umma_f16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

// A: [2][128][64] fp16 (tile of CTA 0, tile of CTA 1); B_image: [2][NH x 128 B] swizzled halves; D: [2][128][N]
__global__ void __launch_bounds__(128, 1) pair_check(const __half* A, const uint8_t* B_image, int N, int dcol, int acol, float* D) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t s_done, s_afull;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int NH = N / 2;
    for (int i = tid; i < NH * 128 / 16; i += 128) ((uint4*)smem)[i] = ((const uint4*)(B_image + (size_t)rank * NH * 128))[i];
    if (tid == 0) { mbar_init(&s_done, 1); mbar_init(&s_afull, 8); }   // 4 warps of each CTA announce their A rows
    fence_async_smem();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const __half* Ar = A + (size_t)rank * 128 * 64;
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t v[8];
        for (int c = 0; c < 8; ++c) {
            const __half lo = Ar[tid * 64 + 16 * ks + 2 * c], hi = Ar[tid * 64 + 16 * ks + 2 * c + 1];
            v[c] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
        }
        tmem_st8(lane_addr + (uint32_t)(acol + 8 * ks), v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(&s_afull, 0);    // on the LEADER's barrier, from both CTAs
    if (rank == 0 && warp == 0) {
        mbar_wait_cluster(&s_afull, 0);
        if (elect_one()) {
            tc_fence_after();
            const uint64_t b = make_desc(smem_u32(smem));
            const uint32_t idesc = make_idesc_m(256, N);
            for (int ks = 0; ks < 4; ++ks) umma_f16_ts_pair(tmem_base + dcol, tmem_base + acol + 8 * ks, b + 2 * ks, idesc, ks ? 1u : 0u);
            umma_commit_pair(&s_done);
        }
        __syncwarp();
    }
    mbar_wait(&s_done, 0);
    tc_fence_after();
    float* Dr = D + (size_t)rank * 128 * N;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + (uint32_t)(dcol + c0), v);
        tmem_ld_wait();
        for (int c = 0; c < 32 && c0 + c < N; ++c) Dr[tid * N + c0 + c] = __uint_as_float(v[c]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)); }
}

static int run(int N, int dcol, int acol) {
    const int NH = N / 2;
    std::vector<__half> A(2 * 128 * 64);
    std::vector<float> Af(2 * 128 * 64), Bf((size_t)N * 64);
    std::vector<uint8_t> img((size_t)N * 128, 0);
    srand(99 + N);
    for (size_t i = 0; i < A.size(); ++i) { Af[i] = (float)(rand() % 17 - 8); A[i] = __float2half(Af[i]); }
    for (int r = 0; r < N; ++r)
        for (int k = 0; k < 64; ++k) {
            Bf[(size_t)r * 64 + k] = (float)(rand() % 13 - 6);
            const __half h = __float2half(Bf[(size_t)r * 64 + k]);
            // row r of B lives in the half of CTA r / NH, local row r % NH
            memcpy(img.data() + (size_t)(r / NH) * NH * 128 + operand_offset(NH, r % NH, k), &h, 2);
        }
    __half* dA; uint8_t* dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, img.size()); cudaMalloc(&dD, sizeof(float) * 2 * 128 * N);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, sizeof(float) * 2 * 128 * N);
    const size_t smem = (size_t)NH * 128 + 1024;
    cudaFuncSetAttribute(pair_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, pair_check, (const __half*)dA, (const uint8_t*)dB, N, dcol, acol, dD);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    std::vector<float> D((size_t)2 * 128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst[2] = {0, 0};
    for (int c = 0; c < 2; ++c)
        for (int i = 0; i < 128; ++i)
            for (int j = 0; j < N; ++j) {
                double ref = 0;
                for (int k = 0; k < 64; ++k) ref += (double)Af[(size_t)c * 128 * 64 + i * 64 + k] * Bf[(size_t)j * 64 + k];
                const double d = fabs(ref - D[(size_t)c * 128 * N + (size_t)i * N + j]);
                if (!(d <= worst[c])) worst[c] = d;
            }
    const bool ok = worst[0] == 0 && worst[1] == 0 && e == cudaSuccess;
    printf("N=%3d  D column %3d  A column %3d : max |D - A B^T| = %g (CTA 0), %g (CTA 1)  %s %s\n", N, dcol, acol, worst[0], worst[1], ok ? "OK" : "MISMATCH",
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return ok ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run(64, 128, 0);
    bad += run(192, 128, 64);
    bad += run(256, 256, 0);
    bad += run(176, 128 + 192, 64);
    printf(bad ? "FAILED\n" : "CTA-pair MMA as assumed\n");
    return bad;
}
