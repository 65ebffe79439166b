// the model's statistics and factorisation on the device.
//
//   * mcr_device_moments: mean vector and sample covariance (ddof = 1) of a T x N matrix of returns - what the reference's
//     offline pipeline computes with pandas (preprocessing/preprocessing_stats.py:18,21: train.mean(), train.cov());
//   * mcr_device_transform: the Cholesky factorisation every reference model runs before simulating
//     (gsl_linalg_cholesky_decomp1: serial.c:36-37, omp.c:47-48, cuda.cu:87-88) plus the upper-triangle convention of
//     mcr_build_transform.  Left-looking, one thread per row, every inner product accumulated left to right with one rounding per
//     multiply and per add - the order of the host routine and of the oracle - so the factor is BITWISE the host's
//     (tests/test_gpu_round2.py::test_device_cholesky_is_bitwise_the_host_one).
//
// Neither is on the hot path at the reference's sizes (N = 15: microseconds on the host); they matter when N grows to real
// constituents (N = 1024: 30 ms on one host core against ~3 ms here) or when the covariance is re-estimated per run.
#include "../../include/mcr_b200.h"
#include "mcr_engine.h"

#include <cmath>
#include <vector>

namespace mcr {

// column means: one warp per column, lanes stride over the rows, fixed-shape tree (deterministic)
__global__ void __launch_bounds__(256) column_mean_kernel(const double* __restrict__ X, int T, int N, double* __restrict__ mu) {
    const int col = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (col >= N) return;
    double s = 0.0;
    for (int t = lane; t < T; t += 32) s += X[(size_t)t * N + col];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) mu[col] = s / (double)T;
}

// Sigma = Xc^T Xc / (T - 1), Xc = X - mu: 32 x 32 output tile per block of 16 x 16 threads (2 x 2 per thread), the row range
// walked in slabs of 32 staged in shared memory
__global__ void __launch_bounds__(256) covariance_kernel(const double* __restrict__ X, const double* __restrict__ mu, int T, int N,
                                                         double* __restrict__ Sigma) {
    __shared__ double sa[32][33], sb[32][33];
    const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
    if (bj > bi) return;   // lower triangle of tiles; mirrored on store
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int t0 = 0; t0 < T; t0 += 32) {
        for (int e = threadIdx.x; e < 32 * 32; e += 256) {
            const int r = e >> 5, c = e & 31, t = t0 + r;
            sa[r][c] = (t < T && bi + c < N) ? X[(size_t)t * N + bi + c] - mu[bi + c] : 0.0;
            sb[r][c] = (t < T && bj + c < N) ? X[(size_t)t * N + bj + c] - mu[bj + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const double a0 = sa[r][ty], a1 = sa[r][ty + 16], b0 = sb[r][tx], b1 = sb[r][tx + 16];
            acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]);
            acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
    const double inv = 1.0 / (double)(T - 1);
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int i = bi + ty + 16 * u, j = bj + tx + 16 * v;
            if (i < N && j < N) {
                const double s = acc[u][v] * inv;
                Sigma[(size_t)i * N + j] = s;
                Sigma[(size_t)j * N + i] = s;
            }
        }
}

// Left-looking Cholesky of the lower triangle, in place, column-major storage (element (i, k) at k * ld + i, so that the threads of
// a warp - consecutive rows i - read consecutive addresses).  One CTA, thread i owns row i; column j:
//   t_i = sum_{k<j} a_jk a_ik   (sequential in k, separate rounding of the product and of the sum),  a_ij -= t_i,
//   then the column is scaled by 1 / sqrt(a_jj), diagonal included - exactly mcr_build_transform / the oracle.
// Row j of the factor (the multipliers a_jk) is staged in shared memory once per column.  status: 0, or 1 + the failing pivot.
__global__ void __launch_bounds__(1024, 1) cholesky_lower_kernel(double* __restrict__ Acm, int N, int ld, int* __restrict__ status) {
    extern __shared__ double s_row[];   // N multipliers of the current column + the pivot's scale
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    for (int j = 0; j < N; ++j) {
        for (int k = threadIdx.x; k < j; k += blockDim.x) s_row[k] = Acm[(size_t)k * ld + j];
        __syncthreads();
        for (int i = j + threadIdx.x; i < N; i += blockDim.x) {
            double t = 0.0;
            for (int k = 0; k < j; ++k) t = __dadd_rn(t, __dmul_rn(s_row[k], Acm[(size_t)k * ld + i]));
            Acm[(size_t)j * ld + i] = __dadd_rn(Acm[(size_t)j * ld + i], -t);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double ajj = Acm[(size_t)j * ld + j];
            if (!(ajj > 0.0) || !isfinite(ajj)) s_fail = 1 + j;
            s_row[N] = 1.0 / sqrt(ajj);   // IEEE division and square root: same bits as the host's
        }
        __syncthreads();
        if (s_fail) break;
        const double inv = s_row[N];
        for (int i = j + threadIdx.x; i < N; i += blockDim.x) Acm[(size_t)j * ld + i] = __dmul_rn(Acm[(size_t)j * ld + i], inv);
        __syncthreads();
    }
    if (threadIdx.x == 0) *status = s_fail;
}

// row-major Sigma -> column-major working copy of its lower triangle: element (i, k), k <= i, at k * N + i
__global__ void __launch_bounds__(256) lower_to_colmajor_kernel(const double* __restrict__ Sigma, int N, int lds, double* __restrict__ Lcm) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)N * N) return;
    const int k = (int)(e / N), i = (int)(e % N);
    Lcm[e] = k <= i ? Sigma[(size_t)i * lds + k] : 0.0;
}

// column-major factor + the original Sigma -> dense row-major transform A (lower: zeros above the diagonal; gsl_full: Sigma there)
__global__ void __launch_bounds__(256) assemble_transform_kernel(const double* __restrict__ Lcm, const double* __restrict__ Sigma, int N, int lds,
                                                                 int gsl_full, double* __restrict__ A) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)N * N) return;
    const int i = (int)(e / N), j = (int)(e % N);
    A[e] = j <= i ? Lcm[(size_t)j * N + i] : (gsl_full ? Sigma[(size_t)j * lds + i] : 0.0);
}

}  // namespace mcr

using namespace mcr;

extern "C" {

int mcr_device_moments(mcr_engine* e, const double* returns, int T, int N, double* mu_out, double* Sigma_out) {
    if (!e || !returns || !mu_out || !Sigma_out) return fail("mcr_device_moments: null argument");
    if (T < 2 || N < 1) return fail("mcr_device_moments: need T >= 2 rows and N >= 1 columns (T=%d, N=%d)", T, N);
    MCR_CUDA(cudaSetDevice(e->device));
    double *dX = nullptr, *dmu = nullptr, *dS = nullptr;
    int rc = -1;
    do {
        if (cudaMalloc(&dX, sizeof(double) * (size_t)T * N) != cudaSuccess || cudaMalloc(&dmu, sizeof(double) * N) != cudaSuccess ||
            cudaMalloc(&dS, sizeof(double) * (size_t)N * N) != cudaSuccess) { fail("mcr_device_moments: device allocation failed"); break; }
        if (cudaMemcpyAsync(dX, returns, sizeof(double) * (size_t)T * N, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) { fail("mcr_device_moments: H2D failed"); break; }
        column_mean_kernel<<<(N + 7) / 8, 256, 0, e->stream>>>(dX, T, N, dmu);
        const dim3 grid((N + 31) / 32, (N + 31) / 32);
        covariance_kernel<<<grid, 256, 0, e->stream>>>(dX, dmu, T, N, dS);
        if (cudaMemcpyAsync(mu_out, dmu, sizeof(double) * N, cudaMemcpyDeviceToHost, e->stream) != cudaSuccess ||
            cudaMemcpyAsync(Sigma_out, dS, sizeof(double) * (size_t)N * N, cudaMemcpyDeviceToHost, e->stream) != cudaSuccess ||
            cudaStreamSynchronize(e->stream) != cudaSuccess) { fail("mcr_device_moments: %s", cudaGetErrorString(cudaGetLastError())); break; }
        rc = 0;
    } while (0);
    cudaFree(dX); cudaFree(dmu); cudaFree(dS);
    return rc;
}

int mcr_device_transform(mcr_engine* e, const double* Sigma, int N, int lds, int mode, double* A_out) {
    if (!e || !Sigma || !A_out || N <= 0 || lds < N) return fail("mcr_device_transform: bad arguments");
    if (mode != MCR_TRANSFORM_LOWER && mode != MCR_TRANSFORM_GSL_FULL) return fail("mcr_device_transform: unknown mode %d", mode);
    if (N > 4096) return fail("mcr_device_transform: N=%d beyond the single-CTA factorisation (4096)", N);
    MCR_CUDA(cudaSetDevice(e->device));
    double *dSig = nullptr, *dL = nullptr, *dA = nullptr;
    int* dstat = nullptr;
    int rc = -1;
    do {
        if (cudaMalloc(&dSig, sizeof(double) * (size_t)N * lds) != cudaSuccess || cudaMalloc(&dL, sizeof(double) * (size_t)N * N) != cudaSuccess ||
            cudaMalloc(&dA, sizeof(double) * (size_t)N * N) != cudaSuccess || cudaMalloc(&dstat, sizeof(int)) != cudaSuccess) {
            fail("mcr_device_transform: device allocation failed");
            break;
        }
        if (cudaMemcpyAsync(dSig, Sigma, sizeof(double) * (size_t)N * lds, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) { fail("mcr_device_transform: H2D failed"); break; }
        // column-major working copy of the lower triangle (like the host routine, only the lower triangle of Sigma is read)
        lower_to_colmajor_kernel<<<(unsigned)(((size_t)N * N + 255) / 256), 256, 0, e->stream>>>(dSig, N, lds, dL);
        const size_t smem = sizeof(double) * (size_t)(N + 1);
        cudaFuncSetAttribute(cholesky_lower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cholesky_lower_kernel<<<1, 1024, smem, e->stream>>>(dL, N, N, dstat);
        assemble_transform_kernel<<<(unsigned)(((size_t)N * N + 255) / 256), 256, 0, e->stream>>>(dL, dSig, N, lds, mode == MCR_TRANSFORM_GSL_FULL, dA);
        int status = 0;
        if (cudaMemcpyAsync(&status, dstat, sizeof(int), cudaMemcpyDeviceToHost, e->stream) != cudaSuccess ||
            cudaMemcpyAsync(A_out, dA, sizeof(double) * (size_t)N * N, cudaMemcpyDeviceToHost, e->stream) != cudaSuccess ||
            cudaStreamSynchronize(e->stream) != cudaSuccess) { fail("mcr_device_transform: %s", cudaGetErrorString(cudaGetLastError())); break; }
        if (status != 0) { fail("mcr_device_transform: Sigma is not positive definite (pivot %d)", status - 1); break; }
        rc = 0;
    } while (0);
    cudaFree(dSig); cudaFree(dL); cudaFree(dA); cudaFree(dstat);
    return rc;
}

}  // extern "C"
