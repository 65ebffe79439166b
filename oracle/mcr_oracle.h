/*
 * CPU oracle for the hot path of javidsegura/gpu-montecarlo-risk — TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C, fp64 restatement of the reference algorithm
 * (src/02-C-serial/monte_carlo_serial.c:21-58,62-85,98-164).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the
 * product path (libmcr_b200.so + the get_cuda_model plugin) never links or
 * calls anything in oracle/.
 *
 * Parity status: the reference has no tests or golden vectors and its
 * arithmetic lives in GSL 2.7, which is not available here.  The oracle is
 * pinned against (1) numpy-legacy MT19937 known-answer vectors, (2) a run of
 * the reference's own Python model (src/01-python-serial/python_serial.py)
 * on identical normals (tests/golden/), and (3) the reference's unmodified C
 * models compiled against the GSL-subset shim (oracle/_ref).  It is UNPINNED
 * against a real libgsl build ("parity unpinned at the GSL boundary").
 */
#ifndef MCR_ORACLE_H
#define MCR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* transform-matrix conventions */
#define MCR_ORACLE_TRANSFORM_LOWER 0    /* A = tril(L)            : reference CUDA model, monte_carlos_cuda.cu:100-107 */
#define MCR_ORACLE_TRANSFORM_GSL_FULL 1 /* A = L + striu(Sigma)   : C-serial/OpenMP after gsl_linalg_cholesky_decomp1 + full dgemv */
#define MCR_ORACLE_TRANSFORM_DIAG_SQRT 2 /* OR-ed in: L_jj = sqrt(a_jj) instead of a_jj * (1 / sqrt(a_jj)) (one-ulp ambiguity of the recalled GSL code) */

/* Sigma (N x N row-major, fp64) -> A (N x N row-major).  Returns 0, or -1 if not SPD. */
int mcr_oracle_transform(const double *sigma, int N, int mode, double *A_out);

/* One trial given its N normals: R = mu + A z in gslcblas order; returns S = #{R_i < -x}. */
int mcr_oracle_trial(const double *A, const double *mu, int N, double x, const double *z, double *R_scratch);

/* M trials, normals drawn like the C-serial model (MT19937 + GSL polar Gaussian, serial.c:46,54,67-69).
 * S_values may be NULL.  Z_out (M*N doubles) may be NULL; when given it receives the normals drawn. */
int mcr_oracle_run_mt19937(const double *A, const double *mu, int N, int k, double x, int64_t M,
                           unsigned long seed, int32_t *S_values, int64_t *count, int64_t *sum_S,
                           double *Z_out);

/* M trials on caller-supplied normals Z (M x N row-major fp64). */
int mcr_oracle_run_normals(const double *A, const double *mu, int N, int k, double x, int64_t M,
                           const double *Z, int32_t *S_values, int64_t *count, int64_t *sum_S);
/* Same, normals given as fp32 (what the GPU generator produces); arithmetic still fp64. */
int mcr_oracle_run_normals_f32(const double *A, const double *mu, int N, int k, double x, int64_t M,
                               const float *Z, int32_t *S_values, int64_t *count, int64_t *sum_S);

/* Extension metrics (north-star A7; no reference code): portfolio return P = w . R per trial. */
int mcr_oracle_portfolio_normals_f32(const double *A, const double *mu, const double *w, int N, int64_t M,
                                     const float *Z, double *P_out);

/* serial.c:152-158 */
void mcr_oracle_stats(int64_t count, int64_t M, double *P_hat, double *std_error, double *ci_lower, double *ci_upper);

/* Philox4x32-10 (Salmon et al. 2011), the counter-based generator the GPU path uses. */
void mcr_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* The GPU path's counter layout: key = (seed_lo, seed_hi), ctr = (path_lo, path_hi, block, 0);
 * block b yields the uniforms for assets 4b..4b+3.  Fills U (M x Npad4 uint32) with Npad4 = 4*ceil(N/4). */
void mcr_oracle_philox_uniforms(uint64_t seed, uint64_t path_offset, int64_t M, int N, uint32_t *U);
/* Box-Muller on those uniforms with libm in fp32 (the GPU uses MUFU approximations, so this
 * agrees to ~1e-6 absolute, not bitwise).  Z: M x N float. */
void mcr_oracle_philox_normals(uint64_t seed, uint64_t path_offset, int64_t M, int N, float *Z);

/* raw MT19937 outputs / GSL-style gaussians, for known-answer tests */
void mcr_oracle_mt19937_raw(unsigned long seed, int n, uint32_t *out);
void mcr_oracle_mt19937_gaussians(unsigned long seed, int n, double *out);

#ifdef __cplusplus
}
#endif
#endif
