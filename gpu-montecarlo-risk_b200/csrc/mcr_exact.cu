// fp64 exact kernel (also the fix-up kernel), finalize, generator dump and FP32-peak calibration.
//
// exact_kernel evaluates trials in fp64 with the reference's summation order
// (gsl_blas_dgemv row-major NoTrans as used at src/02-C-serial/monte_carlo_serial.c:72-73:
//  temp = sum_j z_j * A_ij, sequential and unfused;  R_i = (0 + 1*temp) + mu_i;  hit = R_i < -x).
// It serves (a) as the fix-up for trials the fp32 kernels flagged as too close to the threshold,
// (b) as MCR_KERNEL_EXACT for whole runs.  One warp per trial: lanes own rows, normals of the
// trial sit in shared memory, A is read column-major so a warp's loads coalesce.
#include "mcr_internal.h"
#include "mcr_rng.cuh"
#include "../../include/mcr_b200.h"

#include <algorithm>

namespace mcr {

constexpr int EXACT_WARPS = 4;

__global__ void __launch_bounds__(EXACT_WARPS * 32) exact_kernel(const ExactParams p) {
    extern __shared__ double s_z[];  // EXACT_WARPS x N
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* z = s_z + (size_t)warp * p.N;

    uint64_t n_items = p.M;
    if (p.flag_list) {
        const uint32_t c = *p.flag_count;
        n_items = c < p.flag_cap ? c : p.flag_cap;
    }
    const uint64_t total_warps = (uint64_t)gridDim.x * EXACT_WARPS;
    const double neg_x = -p.x;

    for (uint64_t w = (uint64_t)blockIdx.x * EXACT_WARPS + warp; w < n_items; w += total_warps) {
        const uint64_t i = p.flag_list ? (uint64_t)p.flag_list[w] : w;
        if (p.Z) {
            for (int j = lane; j < p.N; j += 32) z[j] = p.Z[i * (uint64_t)p.N + j];
        } else {
            const uint64_t path = p.path_offset + i;
            for (int b = lane; 4 * b < p.N; b += 32) {
                float zf[4];
                normals4((uint32_t)path, (uint32_t)(path >> 32), (uint32_t)b, p.key0, p.key1, zf);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (4 * b + q < p.N) z[4 * b + q] = (double)zf[q];
            }
        }
        __syncwarp();
        int S = 0;
        // four rows per lane at a time: each row's sum is one dependent fp64 chain (reference order), four
        // independent chains keep the FP64 pipe busy while a chain waits on its previous add
        for (int r0 = lane; r0 < p.N; r0 += 128) {
            double t[4] = {0.0, 0.0, 0.0, 0.0};
            int jend[4], jmax = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = r0 + 32 * q;
                jend[q] = r < p.N ? (p.lower ? r + 1 : p.N) : 0;
                jmax = jend[q] > jmax ? jend[q] : jmax;
            }
            const double* a = p.A_cm + r0;
            for (int j = 0; j < jmax; ++j) {
                const double zj = z[j];
                const double* aj = a + (size_t)j * p.N;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (j < jend[q]) t[q] = __dadd_rn(t[q], __dmul_rn(zj, __ldg(aj + 32 * q)));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = r0 + 32 * q;
                if (r < p.N) {
                    double R = __dadd_rn(0.0, __dmul_rn(1.0, t[q]));
                    R = __dadd_rn(R, p.mu[r]);
                    S += (R < neg_x) ? 1 : 0;
                }
            }
        }
        S = __reduce_add_sync(0xffffffffu, S);
        if (lane == 0) {
            if (S >= p.k) atomicAdd(&p.accum[0], 1ull);
            atomicAdd(&p.accum[1], (unsigned long long)S);
            if (p.S_out) p.S_out[i] = S;
            if (p.hist) atomicAdd(&p.hist[S], 1ull);
        }
        __syncwarp();
    }
}

// Blocked flavour for N > 128.  The warp-per-trial kernel above streams the whole transform (N^2 * 4..8 bytes,
// far more than L1) once per trial and ends up bound by L2 bandwidth.  Here a block takes EXB_T trials at a
// time, keeps their normals in shared memory ([column][trial], so one broadcast LDS.128 feeds two trials)
// and every thread carries two rows x EXB_T independent fp64 chains: each transform element loaded is
// used EXB_T times and the FP64 pipe never waits on a dependent add.  Per row the arithmetic and its order
// are exactly the ones above.
#ifndef MCR_EXB_T
#define MCR_EXB_T 8
#endif
#ifndef MCR_EXB_MINB
#define MCR_EXB_MINB 4
#endif
constexpr int EXB_THREADS = 128, EXB_T = MCR_EXB_T;

__device__ __forceinline__ void exb_step(double (&acc)[EXB_T], const double a, const double* __restrict__ zj) {
#pragma unroll
    for (int q = 0; q < EXB_T; q += 2) {
        const double2 z = *reinterpret_cast<const double2*>(zj + q);
        acc[q] = __dadd_rn(acc[q], __dmul_rn(z.x, a));
        acc[q + 1] = __dadd_rn(acc[q + 1], __dmul_rn(z.y, a));
    }
}

// columns [j0, j1) of ROWS (1 or 2) rows, 128 apart; transform elements are fetched EXB_AHEAD columns at a
// time, one batch ahead of the arithmetic, so the L2 latency of the loads is covered by the batch in flight
constexpr int EXB_AHEAD = 4;
template <int ROWS>
__device__ __forceinline__ void exb_span(double (&acc0)[EXB_T], double (&acc1)[EXB_T], const double* __restrict__ a, int N,
                                         int j0, int j1, const double* __restrict__ s_zt) {
    double cur[ROWS][EXB_AHEAD], nxt[ROWS][EXB_AHEAD];
    const int n_batches = (j1 - j0) / EXB_AHEAD;
    auto fetch = [&](double (&buf)[ROWS][EXB_AHEAD], int j) {
#pragma unroll
        for (int u = 0; u < EXB_AHEAD; ++u) {
            buf[0][u] = __ldg(a + (size_t)(j + u) * N);
            if (ROWS == 2) buf[ROWS - 1][u] = __ldg(a + (size_t)(j + u) * N + EXB_THREADS);
        }
    };
    if (n_batches > 0) fetch(cur, j0);
    int j = j0;
    for (int b = 0; b < n_batches; ++b, j += EXB_AHEAD) {
        if (b + 1 < n_batches) fetch(nxt, j + EXB_AHEAD);
#pragma unroll
        for (int u = 0; u < EXB_AHEAD; ++u) {
            exb_step(acc0, cur[0][u], s_zt + (j + u) * EXB_T);
            if (ROWS == 2) exb_step(acc1, cur[ROWS - 1][u], s_zt + (j + u) * EXB_T);
        }
#pragma unroll
        for (int u = 0; u < EXB_AHEAD; ++u) {
            cur[0][u] = nxt[0][u];
            if (ROWS == 2) cur[ROWS - 1][u] = nxt[ROWS - 1][u];
        }
    }
    for (; j < j1; ++j) {
        exb_step(acc0, __ldg(a + (size_t)j * N), s_zt + j * EXB_T);
        if (ROWS == 2) exb_step(acc1, __ldg(a + (size_t)j * N + EXB_THREADS), s_zt + j * EXB_T);
    }
}

__global__ void __launch_bounds__(EXB_THREADS, MCR_EXB_MINB) exact_block_kernel(const ExactParams p) {
    extern __shared__ __align__(16) double s_zt[];   // [N][EXB_T]
    __shared__ int s_S[EXB_T];
    __shared__ uint64_t s_idx[EXB_T];
    constexpr uint64_t NONE = ~0ull;
    const int tid = threadIdx.x, lane = tid & 31;
    const int N = p.N;

    uint64_t n_items = p.M;
    if (p.flag_list) {
        const uint32_t c = *p.flag_count;
        n_items = c < p.flag_cap ? c : p.flag_cap;
    }
    const uint64_t n_groups = (n_items + EXB_T - 1) / EXB_T;
    const double neg_x = -p.x;

    for (uint64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        if (tid < EXB_T) {
            const uint64_t item = grp * EXB_T + tid;
            s_idx[tid] = item < n_items ? (p.flag_list ? (uint64_t)p.flag_list[item] : item) : NONE;
            s_S[tid] = 0;
        }
        __syncthreads();
        if (p.Z) {
            for (int e = tid; e < EXB_T * N; e += EXB_THREADS) {
                const int q = e / N, j = e - q * N;
                const uint64_t i = s_idx[q];
                s_zt[j * EXB_T + q] = i != NONE ? p.Z[i * (uint64_t)N + j] : 0.0;
            }
        } else {
            const int nb = (N + 3) / 4;
            for (int e = tid; e < EXB_T * nb; e += EXB_THREADS) {
                const int q = e / nb, b = e - q * nb;
                const uint64_t i = s_idx[q];
                float zf[4] = {0.f, 0.f, 0.f, 0.f};
                if (i != NONE) {
                    const uint64_t path = p.path_offset + i;
                    normals4((uint32_t)path, (uint32_t)(path >> 32), (uint32_t)b, p.key0, p.key1, zf);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (4 * b + c < N) s_zt[(4 * b + c) * EXB_T + q] = (double)zf[c];
            }
        }
        __syncthreads();

        int hits[EXB_T];
#pragma unroll
        for (int q = 0; q < EXB_T; ++q) hits[q] = 0;
        for (int rA = tid; rA < N; rA += 2 * EXB_THREADS) {
            const int rB = rA + EXB_THREADS;
            const int jendA = p.lower ? rA + 1 : N;
            const int jendB = rB < N ? (p.lower ? rB + 1 : N) : 0;
            double accA[EXB_T], accB[EXB_T];
#pragma unroll
            for (int q = 0; q < EXB_T; ++q) { accA[q] = 0.0; accB[q] = 0.0; }
            const double* a = p.A_cm + rA;
            if (rB < N) {
                exb_span<2>(accA, accB, a, N, 0, jendA, s_zt);                       // both rows
                exb_span<1>(accB, accA, a + EXB_THREADS, N, jendA, jendB, s_zt);     // columns only row B has
            } else {
                exb_span<1>(accA, accB, a, N, 0, jendA, s_zt);
            }
            const double muA = p.mu[rA], muB = rB < N ? p.mu[rB] : 0.0;
#pragma unroll
            for (int q = 0; q < EXB_T; ++q) {
                double R = __dadd_rn(__dadd_rn(0.0, __dmul_rn(1.0, accA[q])), muA);
                hits[q] += (R < neg_x) ? 1 : 0;
                if (rB < N) {
                    R = __dadd_rn(__dadd_rn(0.0, __dmul_rn(1.0, accB[q])), muB);
                    hits[q] += (R < neg_x) ? 1 : 0;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < EXB_T; ++q) {
            const int h = __reduce_add_sync(0xffffffffu, hits[q]);
            if (lane == 0 && h) atomicAdd(&s_S[q], h);
        }
        __syncthreads();
        if (tid < EXB_T && s_idx[tid] != NONE) {
            const int S = s_S[tid];
            if (S >= p.k) atomicAdd(&p.accum[0], 1ull);
            atomicAdd(&p.accum[1], (unsigned long long)S);
            if (p.S_out) p.S_out[s_idx[tid]] = S;
            if (p.hist) atomicAdd(&p.hist[S], 1ull);
        }
        __syncthreads();
    }
}

#ifndef MCR_EXACT_BLOCK_MIN_N
#define MCR_EXACT_BLOCK_MIN_N 129
#endif

cudaError_t launch_exact(const ExactParams& p, int num_sms, cudaStream_t s) {
    if (p.N >= MCR_EXACT_BLOCK_MIN_N) {
        const size_t smem = sizeof(double) * (size_t)EXB_T * p.N;
        cudaError_t e = cudaFuncSetAttribute(exact_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int bps = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, exact_block_kernel, EXB_THREADS, smem);
        if (e != cudaSuccess) return e;
        int grid = num_sms * std::max(bps, 1);
        if (!p.flag_list) {
            const uint64_t need = (p.M + EXB_T - 1) / EXB_T;
            if (need < (uint64_t)grid) grid = (int)(need ? need : 1);
        }
        exact_block_kernel<<<grid, EXB_THREADS, smem, s>>>(p);
        return cudaGetLastError();
    }
    const size_t smem = sizeof(double) * (size_t)EXACT_WARPS * p.N;
    int grid = num_sms * 8;
    if (!p.flag_list) {
        const uint64_t need = (p.M + EXACT_WARPS - 1) / EXACT_WARPS;
        if (need < (uint64_t)grid) grid = (int)(need ? need : 1);
    }
    exact_kernel<<<grid, EXACT_WARPS * 32, smem, s>>>(p);
    return cudaGetLastError();
}

// ---- single decisions of the slab path ------------------------------------------------------------------------
// One thread per (trial, row): the trial's normals are re-derived from the counter-based stream in column order
// and the row's sum is accumulated exactly as in exact_kernel.  Slab path: a hit is added to the trial's uint16 in
// S_acc (counted later); N <= 128: the entry carries the trial's other hits, so S is final and counted here.
__global__ void __launch_bounds__(128) exact_pairs_kernel(const PairParams p) {
    const uint32_t c = *p.pair_count;
    const uint32_t n = c < p.pair_cap ? c : p.pair_cap;
    const double neg_x = -p.x;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const unsigned long long e = p.pair_list[t];
        const uint64_t i = p.mode == 2 ? e >> 21 : e >> 16;
        const int row = p.mode == 0 ? (int)(e & 0xFFFFu) : p.mode == 1 ? (int)(e & 0xFFu) : (int)(e & 0x3FFu);
        const int S_rest = p.mode == 1 ? (int)((e >> 8) & 0xFFu) : (int)((e >> 10) & 0x7FFu);
        const uint64_t path = p.path_offset + i;
        const int jend = p.lower ? row + 1 : p.N;
        const double* a = p.A_rm + (size_t)row * p.N;
        double acc = 0.0;
        if (p.Z) {   // the caller's normals: the same products in the same order as exact_kernel's row
            const double* z = p.Z + (size_t)i * (size_t)p.N;
            for (int j = 0; j < jend; ++j) acc = __dadd_rn(acc, __dmul_rn(__ldg(z + j), __ldg(a + j)));
        } else {
            for (int b = 0; 4 * b < jend; ++b) {
                float zf[4];
                normals4((uint32_t)path, (uint32_t)(path >> 32), (uint32_t)b, p.key0, p.key1, zf);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (4 * b + q < jend) acc = __dadd_rn(acc, __dmul_rn((double)zf[q], __ldg(a + 4 * b + q)));
            }
        }
        double R = __dadd_rn(0.0, __dmul_rn(1.0, acc));
        R = __dadd_rn(R, p.mu[row]);
        const int hit = R < neg_x ? 1 : 0;
        if (p.mode == 0) {
            if (hit) atomicAdd(reinterpret_cast<unsigned int*>(p.S_acc) + (i >> 1), 1u << (16 * (unsigned)(i & 1)));
        } else {
            const int S = S_rest + hit;
            if (S >= p.k) atomicAdd(&p.accum[0], 1ull);
            atomicAdd(&p.accum[1], (unsigned long long)S);
            if (p.S_out) p.S_out[i] = S;
        }
    }
}

cudaError_t launch_exact_pairs(const PairParams& p, int num_sms, cudaStream_t s) {
    exact_pairs_kernel<<<num_sms * 8, 128, 0, s>>>(p);
    return cudaGetLastError();
}

// ---- cross-block reduction -------------------------------------------------------------------
// out[0] += M, out[1] += count, out[2] += sum_S, out[3] += n_rechecked, out[4] |= status
__global__ void __launch_bounds__(256) finalize_kernel(const FinalizeParams p) {
    __shared__ unsigned long long s_c[8], s_s[8];
    unsigned long long c = 0, s = 0;
    for (int b = threadIdx.x; b < p.n_blocks; b += blockDim.x) {
        c += p.block_partials[2 * (size_t)b];
        s += p.block_partials[2 * (size_t)b + 1];
    }
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_down_sync(0xffffffffu, c, o);
        s += __shfl_down_sync(0xffffffffu, s, o);
    }
    if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = c; s_s[threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        c = 0; s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { c += s_c[w]; s += s_s[w]; }
        uint64_t flagged = 0, status = 0;
        if (p.flag_count) {
            const uint32_t f = *p.flag_count;
            flagged = f < p.flag_cap ? f : p.flag_cap;
            if (f > p.flag_cap) status = 1;  // fix-up list overflow: result incomplete
        }
        if (p.pair_count) {
            const uint32_t f = *p.pair_count;
            flagged += f < p.pair_cap ? f : p.pair_cap;
            if (f > p.pair_cap) status = 1;
        }
        if (p.accum) { c += p.accum[0]; s += p.accum[1]; }
        p.out[0] += p.M;
        p.out[1] += c;
        p.out[2] += s;
        p.out[3] += flagged;
        p.out[4] |= status;
    }
}

cudaError_t launch_finalize(const FinalizeParams& p, cudaStream_t s) {
    finalize_kernel<<<1, 256, 0, s>>>(p);
    return cudaGetLastError();
}

// ---- histogram of S over the per-trial output (extension F3: distribution of the crash count) ----------
__global__ void __launch_bounds__(256) hist_kernel(uint64_t M, const int32_t* S, int nbins, unsigned long long* hist) {
    extern __shared__ unsigned int s_h[];
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) s_h[b] = 0u;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (uint64_t)gridDim.x * blockDim.x) {
        const int v = S[i];
        // most trials share a few values of S: aggregate equal values inside the warp before the shared atomic
        const unsigned peers = __match_any_sync(__activemask(), v);
        if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&s_h[v < nbins ? v : nbins - 1], (unsigned)__popc(peers));
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x)
        if (s_h[b]) atomicAdd(&hist[b], (unsigned long long)s_h[b]);
}

cudaError_t launch_hist(uint64_t M, const int32_t* S, int nbins, unsigned long long* hist, int num_sms, cudaStream_t s) {
    const int grid = (int)std::min<uint64_t>((uint64_t)num_sms * 8, (M + 255) / 256);
    if (grid <= 0) return cudaSuccess;
    hist_kernel<<<grid, 256, sizeof(unsigned int) * nbins, s>>>(M, S, nbins, hist);
    return cudaGetLastError();
}

// ---- generator dump (parity/debug) -----------------------------------------------------------
__global__ void dump_kernel(uint64_t M, uint64_t path_offset, uint32_t key0, uint32_t key1, int N,
                            float* dZ, uint32_t* dU) {
    const int nb = (N + 3) / 4;
    const uint64_t total = M * (uint64_t)nb;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = t / nb;
        const uint32_t b = (uint32_t)(t % nb);
        const uint64_t path = path_offset + i;
        if (dU) {
            const u32x4 u = philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), b, 0u, key0, key1);
            uint32_t* o = dU + (i * nb + b) * 4;
            o[0] = u.x; o[1] = u.y; o[2] = u.z; o[3] = u.w;
        }
        if (dZ) {
            float z[4];
            normals4((uint32_t)path, (uint32_t)(path >> 32), b, key0, key1, z);
            for (int q = 0; q < 4; ++q)
                if ((int)(4 * b + q) < N) dZ[i * (uint64_t)N + 4 * b + q] = z[q];
        }
    }
}

cudaError_t launch_dump(uint64_t M, uint64_t path_offset, uint32_t key0, uint32_t key1, int N,
                        float* dZ, uint32_t* dU, cudaStream_t s) {
    dump_kernel<<<1184, 256, 0, s>>>(M, path_offset, key0, key1, N, dZ, dU);
    return cudaGetLastError();
}

// ---- FP32 pipe calibration: packed FMAs only, 16 independent chains per thread ----------------
__global__ void __launch_bounds__(256) fp32_peak_kernel(int iters, float* sink) {
    uint64_t a[16];
    const float seed = 1.0f + 1e-7f * (float)threadIdx.x;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float v = seed + (float)q;
        a[q] = (uint64_t)__float_as_uint(v) | ((uint64_t)__float_as_uint(v + 0.5f) << 32);
    }
    const float m = 0.999999f, c = 1e-6f;
    const uint64_t mm = (uint64_t)__float_as_uint(m) | ((uint64_t)__float_as_uint(m) << 32);
    const uint64_t cc = (uint64_t)__float_as_uint(c) | ((uint64_t)__float_as_uint(c) << 32);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 16; ++q)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[q]) : "l"(mm), "l"(cc));
    }
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) acc += __uint_as_float((uint32_t)a[q]) + __uint_as_float((uint32_t)(a[q] >> 32));
    if (acc == 12345.678f) sink[0] = acc;  // keep the chains alive
}

cudaError_t launch_fp32_peak(int grid, int iters, float* sink, cudaStream_t s) {
    fp32_peak_kernel<<<grid, 256, 0, s>>>(iters, sink);
    return cudaGetLastError();
}

}  // namespace mcr
