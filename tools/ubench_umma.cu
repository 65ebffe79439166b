// Micro-benchmark (round-2 lead): at what rate does a back-pressured stream of
// tcgen05.mma.kind::f16 (M = 128, K = 16, both operands in shared memory, fp32 accumulate in TMEM) retire
//   * as a function of N (64 / 96 / 128 / 192 / 256),
//   * when consecutive MMAs accumulate into ONE TMEM tile vs. rotate over 2 / 4 tiles,
//   * when they re-read ONE pair of operand tiles vs. rotate over several (as the transform ring does),
//   * while other warps of the CTA hammer shared memory with stores (as the generator warps do)?
// The single-pass kernel measured 109 clocks per M=128,N=96 MMA against the nominal 48 (profiles/r01_wide_n600_after_fix_ncu_summary.txt).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gpu-montecarlo-risk_b200/csrc -o tools/ubench_umma tools/ubench_umma.cu
// Run:   tools/ubench_umma            (one line per configuration: clocks per MMA on SM 0, and the nominal N/2)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "mcr_tc_common.cuh"
using namespace mcr;
using namespace mcr::tc;

constexpr int A_BYTES = 128 * 128;            // 128 rows x 64 fp16, K-major, 128-byte swizzle
constexpr int MAX_OPS = 4;                    // operand tile pairs to rotate over
constexpr int THREADS = 32 + 256;             // warp 0 issues; warps 1..8 optionally store to shared memory

// A operand in TMEM (".ts" form, as cute::SM100_MMA_F16BF16_TS issues it): 128 lanes x 8 columns of packed fp16 pairs per K = 16
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

// TS = 1: the A operand (4 K-steps = 32 columns) sits in TMEM behind the accumulators instead of in shared memory
template <int N, int TS>
__global__ void __launch_bounds__(THREADS, 1) umma_rate(int iters, int nacc, int nops, int stores, long long* clocks, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int B_BYTES = N * 128;
    uint8_t* s_a = smem;                              // [MAX_OPS][A_BYTES]
    uint8_t* s_b = smem + MAX_OPS * A_BYTES;          // [MAX_OPS][B_BYTES]
    uint8_t* s_scratch = s_b + MAX_OPS * B_BYTES;     // 32 KB the store warps write
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < (MAX_OPS * (A_BYTES + B_BYTES)) / 4; i += THREADS) ((uint32_t*)smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (tid == 0) { mbar_init(&s_bar, 1); s_stop = 0; }
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    constexpr uint32_t IDESC = make_idesc(N);

    if (warp == 0) {
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const uint32_t d = tmem_base + (uint32_t)((it % nacc) * N);
                const int op = it % nops;
                const uint64_t a = make_desc(smem_u32(s_a + op * A_BYTES)), b = make_desc(smem_u32(s_b + op * B_BYTES));
                const uint32_t a_tmem = tmem_base + (uint32_t)(nacc * N + 32 * (TS ? op % ((512 - nacc * N) / 32) : 0));
#pragma unroll
                for (int ks = 0; ks < KBLOCK / 16; ++ks) {
                    if (TS) umma_f16_ts(d, a_tmem + 8 * ks, b + 2 * ks, IDESC, 1u);
                    else umma_f16(d, a + 2 * ks, b + 2 * ks, IDESC, 1u);
                }
            }
            umma_commit(&s_bar);
        }
        __syncwarp();
        mbar_wait(&s_bar, 0);
        t1 = clock64();
        tc_fence_after();
        s_stop = 1;
        if (blockIdx.x == 0 && tid == 0) clocks[0] = t1 - t0;
    } else if (stores) {
        // generator-like shared-memory traffic: 16-byte stores, conflict-free, until the MMAs are done
        float acc = 0.f;
        uint4 v = make_uint4(tid, tid, tid, tid);
        uint4* dst = (uint4*)s_scratch + (tid - 32);
        while (!s_stop) {
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q * 256] = v;
            v.x += 1u;
            acc += 1.f;
        }
        if (acc == 1.5f) sink[0] = acc;
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int N, int TS>
void run(int nacc, int nops, int stores) {
    if (nacc * N + (TS ? 32 : 0) > 512) return;
    long long* d_clk; float* d_sink;
    cudaMalloc(&d_clk, 8); cudaMalloc(&d_sink, 4);
    const size_t smem = MAX_OPS * (A_BYTES + N * 128) + 32768 + 1024;
    cudaFuncSetAttribute(umma_rate<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 20000;   // x 4 K-steps
    umma_rate<N, TS><<<148, THREADS, smem>>>(iters / 10, nacc, nops, stores, d_clk, d_sink);
    umma_rate<N, TS><<<148, THREADS, smem>>>(iters, nacc, nops, stores, d_clk, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long clk = 0;
    cudaMemcpy(&clk, d_clk, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d  A in %s  accumulators %d  operand tiles %d  store warps %s : %7.1f clocks per MMA (nominal %d)%s\n", N, TS ? "TMEM" : "smem", nacc, nops,
           stores ? "on " : "off", (double)clk / (4.0 * iters), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d_clk); cudaFree(d_sink);
}

template <int N>
void sweep() {
    for (int stores = 0; stores < 2; ++stores)
        for (int nops : {1, 4})
            for (int nacc : {1, 2, 4}) run<N, 0>(nacc, nops, stores);
    // A in TMEM (not yet run in round 1): expected at the nominal N / 2
    for (int stores = 0; stores < 2; ++stores)
        for (int nacc : {1, 2}) run<N, 1>(nacc, 1, stores);
}

int main() {
    sweep<64>();
    sweep<96>();
    sweep<128>();
    sweep<192>();
    sweep<256>();
    return 0;
}
