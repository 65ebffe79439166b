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
#include "mcr_internal.h"
#include "mcr_rng.cuh"

namespace mcr {

// per-translation-unit constant storage (each instantiation group has its own copy)
constexpr int CONST_PAIRS = 5120;            // 40 KB: NT=100 dense (5000) / NT=128 lower (4160)
static __constant__ float2 cA2[CONST_PAIRS]; // packed transform, column-major row pairs
static __constant__ float2 cC2[MCR_FFMA_MAX_N_INTERNAL / 2];  // c_i = mu_i + x (row pairs)

template <int NT, bool LOWER>
struct Layout {
    static constexpr int H = NT / 2;  // row pairs
    __host__ __device__ static constexpr int m0(int j) { return LOWER ? j / 2 : 0; }
    __host__ __device__ static constexpr int col_off(int j) {
        int s = 0;
        for (int t = 0; t < j; ++t) s += H - m0(t);
        return s;
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
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}

template <int NT, bool LOWER, bool FROM_NORMALS>
__global__ void __launch_bounds__(128) ffma_path_kernel(const FfmaParams p) {
    using L = Layout<NT, LOWER>;
    constexpr int H = L::H;
    constexpr int NB = (NT + 3) / 4;  // Philox blocks per path

    uint32_t cnt = 0, sum_s = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t* cA = reinterpret_cast<const uint64_t*>(cA2);
    const uint64_t* cC = reinterpret_cast<const uint64_t*>(cC2);

#pragma unroll 1
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.M; i += stride) {
        const uint64_t path = p.path_offset + i;
        const uint32_t path_lo = (uint32_t)path, path_hi = (uint32_t)(path >> 32);

        uint64_t acc[H];
#pragma unroll
        for (int m = 0; m < H; ++m) acc[m] = cC[m];

        float zmax = 0.0f;
        const double* zrow = FROM_NORMALS ? p.Z + i * (uint64_t)p.N : nullptr;

#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float z[4];
            if constexpr (FROM_NORMALS) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = 4 * b + q;
                    z[q] = (j < p.N) ? __double2float_rn(zrow[j]) : 0.0f;
                    zmax = fmaxf(zmax, fabsf(z[q]));
                }
            } else {
                normals4(path_lo, path_hi, (uint32_t)b, p.key0, p.key1, z);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = 4 * b + q;
                if (j < NT) {
                    const uint64_t zz = pack2(z[q], z[q]);
#pragma unroll
                    for (int m = L::m0(j); m < H; ++m)
                        acc[m] = ffma2(cA[L::col_off(j) + m - L::m0(j)], zz, acc[m]);
                }
            }
        }

        // threshold + guard band
        uint32_t S = 0;
        float band = 3.0e38f;
#pragma unroll
        for (int m = 0; m < H; ++m) {
            const uint32_t lo = (uint32_t)acc[m], hi = (uint32_t)(acc[m] >> 32);
            S += (lo >> 31) + (hi >> 31);
            band = fminf(band, fminf(fabsf(__uint_as_float(lo)), fabsf(__uint_as_float(hi))));
        }
        const float delta = FROM_NORMALS ? __fmaf_rn(p.delta_rel, zmax, p.delta_abs) : p.delta;
        if (band < delta) {
            const uint32_t slot = atomicAdd(p.flag_count, 1u);
            if (slot < p.flag_cap) p.flag_list[slot] = (uint32_t)i;
        } else {
            cnt += (S >= (uint32_t)p.k) ? 1u : 0u;
            sum_s += S;
            if (p.S_out) p.S_out[i] = (int32_t)S;
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

template <int NT, bool LOWER, bool FROM_NORMALS>
FfmaVariant make_variant() {
    FfmaVariant v{};
    v.NT = NT; v.lower = LOWER; v.normals = FROM_NORMALS;
    v.n_pairs = Layout<NT, LOWER>::n_pairs;
    v.pack = &pack_transform<NT, LOWER>;
    v.upload_A = [](const float* packed, cudaStream_t s) { return upload_A_impl(packed, Layout<NT, LOWER>::n_pairs, s); };
    v.upload_c = [](const float* c, cudaStream_t s) { return upload_c_impl(c, NT, s); };
    v.launch = [](const FfmaParams& p, int grid, int block, cudaStream_t s) {
        ffma_path_kernel<NT, LOWER, FROM_NORMALS><<<grid, block, 0, s>>>(p);
        return cudaGetLastError();
    };
    v.occupancy = [](int block, int* bps) {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, ffma_path_kernel<NT, LOWER, FROM_NORMALS>, block, 0);
    };
    v.symbol_tag = (const void*)&cA2;
    return v;
}

}  // namespace mcr
