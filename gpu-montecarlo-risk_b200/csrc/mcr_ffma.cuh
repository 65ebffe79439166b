// The fused hot-path kernel, FFMA flavour (asset counts up to MCR_FFMA_MAX_N).
//
// One thread = one trial at a time (grid-stride over the path range), everything in registers:
//   Philox4x32-10 -> Box-Muller normals z_j           (A3: serial.c:67-69 / cuda.cu:45-47)
//   acc_i = (mu_i + x) + sum_j A_ij z_j               (A4: serial.c:72-73 / cuda.cu:51-57)
//   S = #{ acc_i < 0 }  == #{ R_i < -x }              (A5: serial.c:76-81 / cuda.cu:60-65)
//   count += [S >= k], sum_S += S                     (A6: serial.c:137-142 / cuda.cu:69-71)
// followed by a warp-shuffle + shared-memory block reduction; a tiny finalize kernel does the
// cross-block step (no atomics on the hot path, deterministic).
//
// B200 mapping.  The correlate is column-oriented: z_j is produced just in time and scattered
// into NT accumulators, so no Z or R vector is ever stored (the reference keeps both in 1 KB of
// local memory per thread, cuda.cu:34-36).  Accumulators are packed in pairs (rows 2m, 2m+1)
// and updated with the Blackwell packed FMA (PTX fma.rn.f32x2 -> SASS FFMA2): one issue slot
// per two FMAs.  The transform comes from constant memory at compile-time offsets (the loops are
// fully unrolled per instantiated NT), which ptxas turns into LDCU.128 uniform loads feeding the
// FFMA2 uniform-register operand: one load instruction per four FMAs and zero vector-register
// bandwidth for A.  In LOWER mode the strict upper triangle is skipped at compile time.
//
// Exactness.  Decisions are taken in fp32, but a trial whose min_i |acc_i| falls inside the
// guard band delta (a rigorous bound on the fp32 evaluation error, computed on the host from
// |A| and |mu + x|) is not counted here: its index is appended to a list and the fp64 fix-up
// kernel (mcr_exact.cu) re-evaluates it in the reference's summation order.  The union is
// bit-identical to evaluating every trial in fp64 on the same normals.
#pragma once
#ifndef MCR_PREFETCH_LINES
#define MCR_PREFETCH_LINES 0
#endif
#ifndef MCR_PIPE_DEPTH
#define MCR_PIPE_DEPTH 8
#endif
#ifndef MCR_ABLATE
#define MCR_ABLATE 0
#endif
#ifndef MCR_MIN_BLOCKS
#define MCR_MIN_BLOCKS 1
#endif
// 1: the transform is copied to shared memory once per block and streamed from there with 16-byte broadcast loads (every lane the
// same address: one wavefront) instead of uniform constant loads.  The constant path sustains about one 16-byte load per 10 cycles
// and SM sub-partition and misses its first-level cache once the transform exceeds ~2 KB; see DESIGN.md section 3.3.
// Measured (round 2, profiles/r02_ffma_smem_transform.log): N = 16 / 32 / 50 are 5-7 % SLOWER from shared memory (1.81 / 4.42 / 9.42 ms
// against 1.69 / 4.19 / 8.92 per 1e8 trials: a 16-byte LDS occupies the SM's one load/store path for 4 cycles whatever the
// broadcast, which is no better than the constant path's rate), N = 100 is 36 % faster (the 40 KB transform thrashes the constant
// caches) - but AUTO runs the tensor kernel from N = 36.  Off by default.
#ifndef MCR_FFMA_SMEM
#define MCR_FFMA_SMEM 0
#endif
#include "mcr_internal.h"
#include "mcr_rng.cuh"

#include <type_traits>
#include <utility>

namespace mcr {

// Compile-time loop: f(std::integral_constant<int, 0>{}), ..., f(std::integral_constant<int, N-1>{}).
// Used instead of `#pragma unroll` for the big loops: their bodies are so large that nvvm's pragma
// unroll budget refuses to unroll them at NT >= 128, which turns the accumulators into local memory.
template <int... I, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, I...>, F&& f) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

// per-translation-unit constant storage (each instantiation group has its own copy)
constexpr int CONST_PAIRS = 5120;            // 40 KB: NT=100 dense (5000) / NT=128 lower (4160)
static __constant__ __align__(16) float2 cA2[CONST_PAIRS + 2]; // packed transform, column-major row pairs
static __constant__ __align__(16) float2 cC2[MCR_FFMA_MAX_N_INTERNAL / 2 + 2];  // c_i = mu_i + x (row pairs)

template <int NT, bool LOWER>
struct Layout {
    static constexpr int H = NT / 2;  // row pairs
    __host__ __device__ static constexpr int m0(int j) { return LOWER ? j / 2 : 0; }
    // pairs stored before column j; every column starts on a 16-byte boundary (even pair index).
    // Closed form (no loop) so that it folds to an immediate at every unrolled use.
    __host__ __device__ static constexpr int up2(int v) { return (v + 1) & ~1; }
    __host__ __device__ static constexpr int g(int q) {  // sum_{u<q} up2(H - u)
        return q * H - q * (q - 1) / 2 + ((H % 2 == 0) ? q / 2 : (q + 1) / 2);
    }
    __host__ __device__ static constexpr int col_off(int j) {
        return LOWER ? 2 * g(j / 2) + ((j & 1) ? up2(H - j / 2) : 0) : j * up2(H);
    }
    static constexpr int n_pairs = col_off(NT);
    static_assert(NT % 2 == 0, "NT must be even");
    static_assert(n_pairs <= CONST_PAIRS, "transform does not fit the constant bank");
};

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// same, but pinned in program order relative to the (volatile) transform loads: this is what keeps
// the software-pipelined load -> use distance intact through nvvm's scheduler
__device__ __forceinline__ uint64_t ffma2_ordered(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// 16 bytes (two row pairs) from constant memory; asm volatile so the load stays where it is written
__device__ __forceinline__ void ldc128(uint64_t const_addr, uint64_t& a, uint64_t& b) {
    asm volatile("ld.const.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(const_addr));
}
// 16 bytes from shared memory (address in the shared window), same contract
__device__ __forceinline__ void lds128(uint32_t smem_addr, uint64_t& a, uint64_t& b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(smem_addr));
}
// prefetch the line holding a (generic) constant-memory address into the uniform/constant cache
__device__ __forceinline__ void prefetch_const(const void* generic_ptr) {
    asm volatile("prefetchu.L1 [%0];" :: "l"(generic_ptr));
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}

template <int NT, bool LOWER, bool FROM_NORMALS, int PPT>
__global__ void __launch_bounds__(128, MCR_MIN_BLOCKS) ffma_path_kernel(const FfmaParams p) {
    using L = Layout<NT, LOWER>;
    constexpr int H = L::H;
    constexpr int NB = (NT + 3) / 4;  // Philox blocks per path
    constexpr int DEPTH = MCR_PIPE_DEPTH;  // 16-byte transform loads in flight ahead of their use
    constexpr int PREFETCH_LINES = MCR_PREFETCH_LINES;  // 64-byte constant-cache lines touched ahead

    uint32_t cnt = 0, sum_s = 0;
    const uint64_t T = (uint64_t)gridDim.x * blockDim.x;
    // .const-space byte addresses (ld.const takes those, not generic pointers)
    const uint64_t cA_const = (uint64_t)__cvta_generic_to_constant(cA2);
    const uint64_t cC_const = (uint64_t)__cvta_generic_to_constant(cC2);
#if MCR_FFMA_SMEM
    __shared__ __align__(16) float2 sA2[L::n_pairs + 2];
    for (int t = threadIdx.x; t < L::n_pairs; t += blockDim.x) sA2[t] = cA2[t];
    __syncthreads();
    const uint32_t sA_addr = (uint32_t)__cvta_generic_to_shared(sA2);
#define MCR_LOAD_A(e_, a_, b_) lds128(sA_addr + 16u * (uint32_t)(e_), a_, b_)
#else
#define MCR_LOAD_A(e_, a_, b_) ldc128(cA_const + 16 * (e_), a_, b_)
#endif

    // One thread carries PPT trials at a time (trial indices i0, i0 + T, ...): every 16-byte uniform
    // load of the transform then feeds 2 * PPT packed FMAs instead of 2.  The uniform constant-load
    // path (LDCU) sustains only about one 16-byte load per 10 cycles per SM sub-partition
    // (measured), so at PPT = 1 it, not the FMA pipe, bounds the kernel.
#pragma unroll 1
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < p.M; i0 += T * PPT) {
        uint64_t idx[PPT];
        bool valid[PPT];
        uint32_t path_lo[PPT], path_hi[PPT];
#pragma unroll
        for (int pp = 0; pp < PPT; ++pp) {
            idx[pp] = i0 + (uint64_t)pp * T;
            valid[pp] = idx[pp] < p.M;
            const uint64_t path = p.path_offset + idx[pp];
            path_lo[pp] = (uint32_t)path;
            path_hi[pp] = (uint32_t)(path >> 32);
        }

        // acc_i starts at c_i = mu_i + x.  All constant loads are volatile asm on purpose: plain
        // constant loads are loop-invariant, so nvvm hoists every one of them out of the trial loop
        // (hundreds of live values) and ptxas has to rematerialise them anyway.
        uint64_t acc[PPT][H + 1];
#pragma unroll
        for (int m = 0; m < H; m += 2) {
            uint64_t c0, c1;
            ldc128(cC_const + 8 * m, c0, c1);
#pragma unroll
            for (int pp = 0; pp < PPT; ++pp) { acc[pp][m] = c0; acc[pp][m + 1] = c1; }
        }

        // pipeline prologue: the first DEPTH couples of the transform stream
        uint64_t pipe[DEPTH][2];
#pragma unroll
        for (int e = 0; e < DEPTH; ++e) MCR_LOAD_A((e < L::n_pairs / 2 ? e : 0), pipe[e][0], pipe[e][1]);

        float zmax[PPT];
#pragma unroll
        for (int pp = 0; pp < PPT; ++pp) zmax[pp] = 0.0f;

        static_for<NB>([&](auto b_c) {
            constexpr int b = decltype(b_c)::value;
            float z[PPT][4];
#pragma unroll
            for (int pp = 0; pp < PPT; ++pp) {
                if constexpr (FROM_NORMALS) {
                    const double* zrow = p.Z + (valid[pp] ? idx[pp] : i0) * (uint64_t)p.N;
                    if ((p.N & 1) == 0 && (reinterpret_cast<uintptr_t>(p.Z) & 15) == 0 && 4 * b + 4 <= p.N) {
                        // even N: every row starts 16-byte aligned, so the four normals of a block are two 16-byte loads
                        // (the lanes of a warp read rows N * 8 bytes apart: the load unit, not HBM, bounds this mode)
                        const double2 a = __ldg(reinterpret_cast<const double2*>(zrow + 4 * b));
                        const double2 c = __ldg(reinterpret_cast<const double2*>(zrow + 4 * b + 2));
                        z[pp][0] = __double2float_rn(a.x); z[pp][1] = __double2float_rn(a.y);
                        z[pp][2] = __double2float_rn(c.x); z[pp][3] = __double2float_rn(c.y);
#pragma unroll
                        for (int q = 0; q < 4; ++q) zmax[pp] = fmaxf(zmax[pp], fabsf(z[pp][q]));
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int j = 4 * b + q;
                            z[pp][q] = (j < p.N) ? __double2float_rn(zrow[j]) : 0.0f;
                            zmax[pp] = fmaxf(zmax[pp], fabsf(z[pp][q]));
                        }
                    }
                } else {
#if MCR_ABLATE == 1   // tuning only: no RNG, normals replaced by a 2-instruction hash
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        z[pp][q] = __uint_as_float(0x3f800000u | ((path_lo[pp] * 2654435761u + (uint32_t)(4 * b + q) * 40503u) >> 9)) - 1.5f;
#else
                    normals4(path_lo[pp], path_hi[pp], (uint32_t)b, p.key0, p.key1, z[pp]);
#endif
                }
            }
            static_for<4>([&](auto q_c) {
                constexpr int q = decltype(q_c)::value;
                constexpr int j = 4 * b + q;
                if constexpr (j < NT) {
                    uint64_t zz[PPT];
#pragma unroll
                    for (int pp = 0; pp < PPT; ++pp) zz[pp] = pack2(z[pp][q], z[pp][q]);
#if MCR_ABLATE == 2   // tuning only: no correlate, one packed FMA per normal keeps the RNG alive
#pragma unroll
                    for (int pp = 0; pp < PPT; ++pp) acc[pp][j % H] = ffma2(acc[pp][(j + 1) % H], zz[pp], acc[pp][j % H]);
#else
                    constexpr int M0 = L::m0(j);
                    static_for<(H - M0 + 1) / 2>([&](auto t_c) {
                        constexpr int m = M0 + 2 * decltype(t_c)::value;
                        // couple e = rows (2m..2m+3) of column j; the stream of couples is consumed in
                        // storage order through a DEPTH-deep software pipeline of uniform registers
                        constexpr int e = (L::col_off(j) + m - M0) / 2;
                        const uint64_t a0 = pipe[e % DEPTH][0], a1 = pipe[e % DEPTH][1];
#pragma unroll
                        for (int pp = 0; pp < PPT; ++pp) {
                            acc[pp][m] = ffma2_ordered(a0, zz[pp], acc[pp][m]);
                            if constexpr (m + 1 < H) acc[pp][m + 1] = ffma2_ordered(a1, zz[pp], acc[pp][m + 1]);
                        }
                        if constexpr (e + DEPTH < L::n_pairs / 2)
                            MCR_LOAD_A((e + DEPTH), pipe[e % DEPTH][0], pipe[e % DEPTH][1]);
                        // The SM's first-level constant cache is far smaller than the transform (5 KB at
                        // N = 50), so the stream misses to the GPC-level constant cache once per line and
                        // the (loosely lock-stepped) warps of the SM all wait on that miss.  Optionally touch
                        // the line PREFETCH_LINES ahead once per 64-byte line.
                        if constexpr (PREFETCH_LINES > 0 && e % 4 == 0 && 16 * e + 64 * PREFETCH_LINES < 8 * L::n_pairs)
                            prefetch_const(reinterpret_cast<const char*>(cA2) + 16 * e + 64 * PREFETCH_LINES);
                    });
#endif
                }
            });
        });

        // threshold + guard band
#pragma unroll
        for (int pp = 0; pp < PPT; ++pp) {
            uint32_t S = 0;
            float band = 3.0e38f;
#pragma unroll
            for (int m = 0; m < H; ++m) {
                const uint32_t lo = (uint32_t)acc[pp][m], hi = (uint32_t)(acc[pp][m] >> 32);
                S += (lo >> 31) + (hi >> 31);
                band = fminf(band, fminf(fabsf(__uint_as_float(lo)), fabsf(__uint_as_float(hi))));
            }
            const float delta = FROM_NORMALS ? __fmaf_rn(p.delta_rel, zmax[pp], p.delta_abs) : p.delta;
            if (valid[pp]) {
                if (band < delta) {
                    const uint32_t slot = atomicAdd(p.flag_count, 1u);
                    if (slot < p.flag_cap) p.flag_list[slot] = (uint32_t)idx[pp];
                } else {
                    cnt += (S >= (uint32_t)p.k) ? 1u : 0u;
                    sum_s += S;
                    if (p.S_out) p.S_out[idx[pp]] = (int32_t)S;
                }
            }
        }
    }

    // block reduction: warp shuffle (REDUX) then shared memory across the 4 warps
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    sum_s = __reduce_add_sync(0xffffffffu, sum_s);
    __shared__ uint32_t s_cnt[4], s_sum[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_cnt[warp] = cnt; s_sum[warp] = sum_s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t c = 0, s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { c += s_cnt[w]; s += s_sum[w]; }
        p.block_partials[2 * (uint64_t)blockIdx.x] = c;
        p.block_partials[2 * (uint64_t)blockIdx.x + 1] = s;
    }
}

// ---- host side of one instantiation ---------------------------------------------------------
template <int NT, bool LOWER>
void pack_transform(const double* A, int N, float* packed) {
    using L = Layout<NT, LOWER>;
    for (int t = 0; t < 2 * L::n_pairs; ++t) packed[t] = 0.0f;
    for (int j = 0; j < NT; ++j)
        for (int m = L::m0(j); m < L::H; ++m) {
            const int idx = L::col_off(j) + m - L::m0(j);
            for (int h = 0; h < 2; ++h) {
                const int i = 2 * m + h;
                double a = (i < N && j < N) ? A[(size_t)i * N + j] : 0.0;
                if (LOWER && j > i) a = 0.0;
                packed[2 * idx + h] = (float)a;
            }
        }
}

static cudaError_t upload_A_impl(const float* packed, int n_pairs, cudaStream_t s) {
    return cudaMemcpyToSymbolAsync(cA2, packed, sizeof(float2) * (size_t)n_pairs, 0, cudaMemcpyHostToDevice, s);
}
static cudaError_t upload_c_impl(const float* c, int NT, cudaStream_t s) {
    return cudaMemcpyToSymbolAsync(cC2, c, sizeof(float) * (size_t)NT, 0, cudaMemcpyHostToDevice, s);
}

template <int NT, bool LOWER, bool FROM_NORMALS, int PPT>
FfmaVariant make_variant() {
    FfmaVariant v{};
    v.NT = NT; v.lower = LOWER; v.normals = FROM_NORMALS; v.ppt = PPT;
    v.n_pairs = Layout<NT, LOWER>::n_pairs;
    v.pack = &pack_transform<NT, LOWER>;
    v.upload_A = [](const float* packed, cudaStream_t s) { return upload_A_impl(packed, Layout<NT, LOWER>::n_pairs, s); };
    v.upload_c = [](const float* c, cudaStream_t s) { return upload_c_impl(c, NT, s); };
    v.launch = [](const FfmaParams& p, int grid, int block, cudaStream_t s) {
        ffma_path_kernel<NT, LOWER, FROM_NORMALS, PPT><<<grid, block, 0, s>>>(p);
        return cudaGetLastError();
    };
    v.occupancy = [](int block, int* bps) {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, ffma_path_kernel<NT, LOWER, FROM_NORMALS, PPT>, block, 0);
    };
    v.symbol_tag = (const void*)&cA2;
    return v;
}

}  // namespace mcr
