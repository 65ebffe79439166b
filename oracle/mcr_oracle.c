/*
 * CPU oracle — TEST INFRASTRUCTURE ONLY (see mcr_oracle.h for the rules and parity status).
 * Each function cites the reference lines it follows (paths relative to the reference root).
 */
#include "mcr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ----------------------------------------------------------------------------------------
 * MT19937 + polar Gaussian — what `gsl_rng_alloc(gsl_rng_mt19937); gsl_rng_set(rng, seed);
 * gsl_ran_gaussian(rng, 1.0)` computes (src/02-C-serial/monte_carlo_serial.c:46,54,67-69).
 * GSL 2.7 rng/mt.c and randist/gauss.c (third-party, absent here) restated.
 * -------------------------------------------------------------------------------------- */
typedef struct { uint32_t mt[624]; int idx; } mt_t;

static void mt_seed(mt_t *g, unsigned long seed) {
    if (seed == 0) seed = 4357;
    g->mt[0] = (uint32_t)(seed & 0xffffffffUL);
    for (int i = 1; i < 624; i++)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static uint32_t mt_next(mt_t *g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

static double mt_uniform_pos(mt_t *g) { /* gsl_rng_uniform_pos: get()/2^32, reject 0 */
    double u;
    do { u = (double)mt_next(g) / 4294967296.0; } while (u == 0.0);
    return u;
}

static double mt_gaussian(mt_t *g) { /* gsl_ran_gaussian(r, 1.0): polar method, second variate discarded */
    double x, y, r2;
    do {
        x = -1 + 2 * mt_uniform_pos(g);
        y = -1 + 2 * mt_uniform_pos(g);
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return 1.0 * y * sqrt(-2.0 * log(r2) / r2);
}

void mcr_oracle_mt19937_raw(unsigned long seed, int n, uint32_t *out) {
    mt_t g; mt_seed(&g, seed);
    for (int i = 0; i < n; i++) out[i] = mt_next(&g);
}

void mcr_oracle_mt19937_gaussians(unsigned long seed, int n, double *out) {
    mt_t g; mt_seed(&g, seed);
    for (int i = 0; i < n; i++) out[i] = mt_gaussian(&g);
}

/* ----------------------------------------------------------------------------------------
 * Transform matrix: gsl_matrix_memcpy(L, Sigma); gsl_linalg_cholesky_decomp1(L)
 * (monte_carlo_serial.c:36-37; monte_carlo_omp.c:47-48; monte_carlos_cuda.cu:87-88).
 * fp64 Cholesky (unblocked, left-looking — GSL's order for N <= 24); the two conventions
 * for the strict upper triangle are SURVEY.md §0 fact 5.
 * -------------------------------------------------------------------------------------- */
int mcr_oracle_transform(const double *sigma, int N, int mode, double *A) {
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) A[i * N + j] = sigma[i * N + j];
    for (int j = 0; j < N; j++) {
        for (int i = j; i < N; i++) {
            double t = 0.0;
            for (int k = 0; k < j; k++) t += A[j * N + k] * A[i * N + k];
            A[i * N + j] += -1.0 * t;
        }
        double ajj = A[j * N + j];
        if (!(ajj > 0.0)) return -1;
        /* GSL 2.7 cholesky_decomp_L2 ends the column with gsl_vector_scale(A(j:n,j), 1/sqrt(ajj)), diagonal included
         * [recalled; GSL is not installed here], so L_jj = ajj * (1/sqrt(ajj)).  MCR_ORACLE_TRANSFORM_DIAG_SQRT stores
         * L_jj = sqrt(ajj) instead (at most one ulp away); tests/test_oracle_golden.py shows the frozen counts do not
         * depend on which of the two a GSL build uses. */
        const double root = sqrt(ajj), inv = 1.0 / root;
        for (int i = j; i < N; i++) A[i * N + j] *= inv;
        if (mode & MCR_ORACLE_TRANSFORM_DIAG_SQRT) A[j * N + j] = root;
    }
    for (int i = 0; i < N; i++)
        for (int j = i + 1; j < N; j++)
            A[i * N + j] = ((mode & 1) == MCR_ORACLE_TRANSFORM_GSL_FULL) ? sigma[j * N + i] : 0.0;
    return 0;
}

/* ----------------------------------------------------------------------------------------
 * One trial: serial_run_trial (monte_carlo_serial.c:62-85).
 *   gsl_blas_dgemv(CblasNoTrans, 1.0, L, Z, 0.0, R)  -> y=0; temp=sum_j z[j]*A[i][j]; y[i] += 1.0*temp
 *   gsl_vector_add(R, mu)                            -> R[i] += mu[i]
 *   S = #{ R[i] < -x }                               -> strict, fp64
 * Products and sums are separate roundings (no FMA contraction: built with -ffp-contract=off).
 * -------------------------------------------------------------------------------------- */
int mcr_oracle_trial(const double *A, const double *mu, int N, double x, const double *z, double *R) {
    int S = 0;
    for (int i = 0; i < N; i++) {
        double temp = 0.0;
        const double *row = A + (size_t)i * N;
        for (int j = 0; j < N; j++) temp += z[j] * row[j];
        double r = 0.0;
        r += 1.0 * temp;
        r += mu[i];
        if (R) R[i] = r;
        if (r < -x) S++;
    }
    return S;
}

/* serial_simulate main loop (monte_carlo_serial.c:131-143) */
int mcr_oracle_run_mt19937(const double *A, const double *mu, int N, int k, double x, int64_t M,
                           unsigned long seed, int32_t *S_values, int64_t *count, int64_t *sum_S,
                           double *Z_out) {
    mt_t g; mt_seed(&g, seed);
    double *z = (double *)malloc(sizeof(double) * (size_t)N);
    if (!z) return -1;
    int64_t c = 0, s = 0;
    for (int64_t j = 0; j < M; j++) {
        for (int i = 0; i < N; i++) z[i] = mt_gaussian(&g);
        if (Z_out) memcpy(Z_out + (size_t)j * N, z, sizeof(double) * (size_t)N);
        int S = mcr_oracle_trial(A, mu, N, x, z, NULL);
        if (S_values) S_values[j] = S;
        s += S;
        if (S >= k) c++;
    }
    free(z);
    *count = c;
    if (sum_S) *sum_S = s;
    return 0;
}

int mcr_oracle_run_normals(const double *A, const double *mu, int N, int k, double x, int64_t M,
                           const double *Z, int32_t *S_values, int64_t *count, int64_t *sum_S) {
    int64_t c = 0, s = 0;
    for (int64_t j = 0; j < M; j++) {
        int S = mcr_oracle_trial(A, mu, N, x, Z + (size_t)j * N, NULL);
        if (S_values) S_values[j] = S;
        s += S;
        if (S >= k) c++;
    }
    *count = c;
    if (sum_S) *sum_S = s;
    return 0;
}

int mcr_oracle_run_normals_f32(const double *A, const double *mu, int N, int k, double x, int64_t M,
                               const float *Z, int32_t *S_values, int64_t *count, int64_t *sum_S) {
    double *z = (double *)malloc(sizeof(double) * (size_t)N);
    if (!z) return -1;
    int64_t c = 0, s = 0;
    for (int64_t j = 0; j < M; j++) {
        for (int i = 0; i < N; i++) z[i] = (double)Z[(size_t)j * N + i];
        int S = mcr_oracle_trial(A, mu, N, x, z, NULL);
        if (S_values) S_values[j] = S;
        s += S;
        if (S >= k) c++;
    }
    free(z);
    *count = c;
    if (sum_S) *sum_S = s;
    return 0;
}

int mcr_oracle_portfolio_normals_f32(const double *A, const double *mu, const double *w, int N, int64_t M,
                                     const float *Z, double *P_out) {
    double *z = (double *)malloc(sizeof(double) * (size_t)N * 2);
    if (!z) return -1;
    double *R = z + N;
    for (int64_t j = 0; j < M; j++) {
        for (int i = 0; i < N; i++) z[i] = (double)Z[(size_t)j * N + i];
        mcr_oracle_trial(A, mu, N, 0.0, z, R);
        double p = 0.0;
        for (int i = 0; i < N; i++) p += w[i] * R[i];
        P_out[j] = p;
    }
    free(z);
    return 0;
}

/* monte_carlo_serial.c:152-158 */
void mcr_oracle_stats(int64_t count, int64_t M, double *P_hat, double *std_error, double *ci_lower, double *ci_upper) {
    double p = (double)count / (double)M;
    double se = sqrt(p * (1.0 - p) / (double)M);
    double margin = 1.96 * se;
    *P_hat = p;
    *std_error = se;
    *ci_lower = fmax(0.0, p - margin);
    *ci_upper = fmin(1.0, p + margin);
}

/* ----------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11).  No reference counterpart (the reference GPU model uses cuRAND XORWOW,
 * monte_carlos_cuda.cu:30-31,45-47); this restates the published algorithm so the GPU
 * generator's integer stream can be checked bit-for-bit.
 * -------------------------------------------------------------------------------------- */
void mcr_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void mcr_oracle_philox_uniforms(uint64_t seed, uint64_t path_offset, int64_t M, int N, uint32_t *U) {
    const int nb = (N + 3) / 4;
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int64_t p = 0; p < M; p++) {
        const uint64_t path = path_offset + (uint64_t)p;
        for (int b = 0; b < nb; b++) {
            const uint32_t ctr[4] = {(uint32_t)path, (uint32_t)(path >> 32), (uint32_t)b, 0u};
            mcr_oracle_philox4x32_10(ctr, key, U + ((size_t)p * nb + b) * 4);
        }
    }
}

/* Box-Muller as the GPU path defines it (gpu-montecarlo-risk_b200/csrc/mcr_rng.cuh):
 *   u_r = u0 * 2^-32 + 2^-33 in (0, 1];  rad = sqrt(-2 ln u_r);  th = 2 pi * u1 * 2^-32
 *   z0 = rad cos(th), z1 = rad sin(th); (u2, u3) -> (z2, z3) likewise. */
void mcr_oracle_philox_normals(uint64_t seed, uint64_t path_offset, int64_t M, int N, float *Z) {
    const int nb = (N + 3) / 4;
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int64_t p = 0; p < M; p++) {
        const uint64_t path = path_offset + (uint64_t)p;
        for (int b = 0; b < nb; b++) {
            const uint32_t ctr[4] = {(uint32_t)path, (uint32_t)(path >> 32), (uint32_t)b, 0u};
            uint32_t u[4];
            float z[4];
            mcr_oracle_philox4x32_10(ctr, key, u);
            for (int h = 0; h < 2; h++) {
                float ur = fmaf((float)u[2 * h], 2.3283064365386963e-10f, 1.1641532182693481e-10f);
                float rad = sqrtf(-2.0f * logf(ur));
                float th = (float)u[2 * h + 1] * (6.2831853071795865f * 2.3283064365386963e-10f);
                z[2 * h] = rad * cosf(th);
                z[2 * h + 1] = rad * sinf(th);
            }
            for (int q = 0; q < 4; q++)
                if (4 * b + q < N) Z[(size_t)p * N + 4 * b + q] = z[q];
        }
    }
}
