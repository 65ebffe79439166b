/*
 * mcr_cli — run the B200 hot path straight from the preprocessing binaries, no GSL / libyaml needed.
 *
 *   mcr_cli <data_dir> [M] [x] [k] [seed] [n_gpus] [lower|gsl_full]
 *
 * <data_dir> holds mu.bin / sigma.bin as written by the reference's preprocessing/save_binary.py:23-45.
 * This is the loader + transform upload + run sequence of main_runner.c:113-130,248 reduced to the hot
 * path: model_loader.c (hardened twin of src/utilities/load_binary.c) -> mcr_build_transform ->
 * mcr_simulate_host.  M is 64-bit here (the reference's int M stops at 2.1e9).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcr_b200.h"
#include "utilities/model_loader.h"

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <data_dir> [M=1000000] [x=0.02] [k=5] [seed=67] [n_gpus=1] [lower|gsl_full]\n", argv[0]);
        return 2;
    }
    const char *dir = argv[1];
    const unsigned long long M = argc > 2 ? strtoull(argv[2], NULL, 10) : 1000000ULL;
    const double x = argc > 3 ? atof(argv[3]) : 0.02;
    const int k = argc > 4 ? atoi(argv[4]) : 5;
    const unsigned long long seed = argc > 5 ? strtoull(argv[5], NULL, 10) : 67ULL;
    const int n_gpus = argc > 6 ? atoi(argv[6]) : 1;
    const int mode = (argc > 7 && strcmp(argv[7], "gsl_full") == 0) ? MCR_TRANSFORM_GSL_FULL : MCR_TRANSFORM_LOWER;

    char mu_path[4096], sigma_path[4096];
    snprintf(mu_path, sizeof mu_path, "%s/mu.bin", dir);
    snprintf(sigma_path, sizeof sigma_path, "%s/sigma.bin", dir);
    /* the payload is staged in pinned host memory (cudaHostAlloc behind the C ABI); header, length and dimension checks come
     * first, so a bad file is rejected before any allocation (and without a GPU) */
    McrHostModel m;
    const McrHostAllocator pinned = {mcr_host_alloc_pinned, mcr_host_free_pinned};
    if (mcr_load_host_model_ex(mu_path, sigma_path, &pinned, &m) != 0) return 1;

    double *A = (double *)malloc((size_t)m.N * (size_t)m.N * sizeof(double));
    if (!A || mcr_build_transform(m.Sigma, m.N, m.N, mode, A) != 0) {
        fprintf(stderr, "Error: %s\n", A ? mcr_last_error() : "out of memory");
        mcr_free_host_model(&m);
        free(A);
        return 1;
    }
    mcr_host_request req;
    memset(&req, 0, sizeof req);
    req.struct_size = sizeof req;
    req.N = m.N; req.k = k; req.M = M; req.seed = seed; req.x = x;
    req.mu = m.mu; req.A = A; req.a_is_lower = (mode == MCR_TRANSFORM_LOWER);
    req.n_gpus = n_gpus;
    req.opt.struct_size = sizeof req.opt;
    mcr_result *r = (mcr_result *)calloc(1, sizeof(mcr_result));
    int rc = r ? mcr_simulate_host(&req, r) : -1;
    if (rc != 0) {
        fprintf(stderr, "Error: %s\n", r ? mcr_last_error() : "out of memory");
    } else {
        printf("N=%d M=%llu x=%g k=%d seed=%llu gpus=%d transform=%s\n", m.N, M, x, k, seed, n_gpus,
               mode == MCR_TRANSFORM_LOWER ? "lower" : "gsl_full");
        printf("count=%llu sum_S=%llu P_hat=%.9f std_error=%.9f ci=[%.9f, %.9f]\n", (unsigned long long)r->count,
               (unsigned long long)r->sum_S, r->P_hat, r->std_error, r->ci_lower, r->ci_upper);
        printf("kernel=%d padded_N=%d launches=%d rechecked=%llu device_ms=%.3f sims_per_s=%.4e\n", r->kernel_used, r->padded_N,
               r->launches, (unsigned long long)r->n_rechecked, r->device_ms,
               r->device_ms > 0 ? (double)r->n_paths / (r->device_ms * 1e-3) : 0.0);
    }
    free(r); free(A);
    mcr_free_host_model(&m);
    return rc == 0 ? 0 : 1;
}
