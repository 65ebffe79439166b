/*
 * B200 model plugin: `get_cuda_model()` behind the reference's src/model_interface.h.
 *
 * Drop-in for src/05-GPU/monte_carlos_cuda.cu (reference): main_runner.c:24-26,230-237 selects it
 * when config.yaml lists `cuda` (or `gpu`) and the driver is built with -DCUDA_BUILD.  Same contract
 * as the other models (main_runner.c:248-295): fills P_hat/count/std_error/ci_*, returns 0 / -1,
 * prints to stdout, errors to stderr; result->S_values is malloc-owned (the driver free()s it) or
 * NULL (make_clean is NULL-safe).
 *
 * Differences from the reference CUDA model, by design:
 *   - no N <= 64 cap (cuda.cu:78-81): FFMA2 kernel for small N, tcgen05 tensor-core kernels above (one tile kernel up to
 *     N = 128, the streamed multi-pass kernels up to N = 1024), every near-threshold decision rechecked in fp64;
 *   - S_values (M ints, written by every reference model but read by nobody) is only produced when
 *     MCR_S_VALUES=1, which saves a 4*M-byte device buffer and D2H copy;
 *   - the transform convention is selectable (MCR_TRANSFORM=lower|gsl_full, see cholesky_upload.h);
 *   - paths can be sharded over MCR_NUM_GPUS devices with one NCCL all-reduce;
 *   - MCR_TIMING_CSV=<path> appends the device-timed ms / sims/s of every run to a sibling CSV (the driver's
 *     own results row keeps its wall-clock execution_time_ms, csv_writer.c:41-45);
 *   - all CUDA work happens in libmcr_b200.so through its C ABI (include/mcr_b200.h); this file is
 *     plain C and has no CPU compute fallback.
 * No config struct reaches a model (model_interface.h:30), hence the environment variables.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "model_interface.h"
#include "mcr_b200.h"
#include "utilities/cholesky_upload.h"

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

static int b200_simulate(MonteCarloParams *params, MonteCarloResult *result) {
    printf("Starting CUDA (B200, sm_100a) Monte Carlo simulation with M = %d trials...\n", params->M);
    printf("Parameters: N=%d, k=%d, x=%.2f%%\n", params->N, params->k, params->x * 100);
    if (params->M < 0 || params->N <= 0 || !params->mu || !params->Sigma) {
        fprintf(stderr, "Error: invalid MonteCarloParams\n");
        return -1;
    }
    if ((size_t)params->N != params->mu->size) {
        fprintf(stderr, "Error: params->N=%d but mu has %zu entries\n", params->N, params->mu->size);
        return -1;
    }

    const int mode = mcr_transform_mode_from_env();
    double *mu = NULL, *A = NULL;
    if (mcr_prepare_model(params->mu, params->Sigma, mode, &mu, &A) != 0) return -1;

    mcr_host_request req;
    memset(&req, 0, sizeof(req));
    req.struct_size = sizeof(req);
    req.N = params->N;
    req.k = params->k;
    req.M = (uint64_t)params->M;
    req.seed = (uint64_t)params->random_seed;
    req.x = params->x;
    req.mu = mu;
    req.A = A;
    req.a_is_lower = (mode == MCR_TRANSFORM_LOWER);
    req.n_gpus = env_int("MCR_NUM_GPUS", 1);
    req.first_device = env_int("MCR_FIRST_DEVICE", 0);
    req.opt.struct_size = sizeof(req.opt);
    req.opt.kernel = MCR_KERNEL_AUTO;

    result->count = 0;
    result->S_values = NULL;
    if (env_int("MCR_S_VALUES", 0) && params->M > 0) {
        result->S_values = (int *)malloc((size_t)params->M * sizeof(int));
        if (!result->S_values) {
            fprintf(stderr, "Error: Failed to allocate S_values array\n");
            free(mu); free(A);
            return -1;
        }
        req.want_S_values = 1;
        req.S_values_out = (int32_t *)result->S_values;
    }

    mcr_result r;
    const int rc = mcr_simulate_host(&req, &r);
    free(mu);
    free(A);
    if (rc != 0) {
        fprintf(stderr, "Error: B200 simulation failed: %s\n", mcr_last_error());
        free(result->S_values);
        result->S_values = NULL;
        return -1;
    }
    if (r.count > (uint64_t)INT_MAX) { /* cannot happen while M is an int, kept as a guard */
        fprintf(stderr, "Error: count %llu does not fit MonteCarloResult.count\n", (unsigned long long)r.count);
        free(result->S_values);
        result->S_values = NULL;
        return -1;
    }
    result->count = (int)r.count;
    result->P_hat = r.P_hat;
    result->std_error = r.std_error;
    result->ci_lower = r.ci_lower;
    result->ci_upper = r.ci_upper;
    /* F2: device-side timing next to the driver's wall-clock row, without touching its 22-column schema:
     * MCR_TIMING_CSV=<path> appends one line per run (header written once). */
    {
        const char *tpath = getenv("MCR_TIMING_CSV");
        if (tpath && *tpath) {
            FILE *probe = fopen(tpath, "r");
            FILE *f = fopen(tpath, "a");
            if (f) {
                if (!probe) fprintf(f, "model_name,N,M,k,x,seed,n_gpus,kernel,padded_N,launches,device_ms,sims_per_s,rechecked,count,P_hat\n");
                fprintf(f, "cuda,%d,%d,%d,%.6f,%lu,%d,%d,%d,%d,%.4f,%.6e,%llu,%llu,%.9f\n", params->N, params->M, params->k, params->x,
                        params->random_seed, req.n_gpus > 0 ? req.n_gpus : 1, r.kernel_used, r.padded_N, r.launches, r.device_ms,
                        r.device_ms > 0 ? (double)r.n_paths / (r.device_ms * 1e-3) : 0.0, (unsigned long long)r.n_rechecked,
                        (unsigned long long)r.count, r.P_hat);
                fclose(f);
            }
            if (probe) fclose(probe);
        }
    }
    printf("CUDA Simulation complete (kernel=%d, padded N=%d, %d launches, device time %.3f ms, %.3e simulations/s, "
           "%llu trials rechecked in fp64)\n",
           r.kernel_used, r.padded_N, r.launches, r.device_ms,
           r.device_ms > 0 ? (double)r.n_paths / (r.device_ms * 1e-3) : 0.0, (unsigned long long)r.n_rechecked);
    return 0;
}

ModelFunctions get_cuda_model(void) {
    ModelFunctions model = {
        .name = "cuda",
        .run_model = b200_simulate
    };
    return model;
}
