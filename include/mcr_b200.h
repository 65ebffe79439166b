/*
 * mcr_b200.h — C ABI of libmcr_b200.so: the B200 (sm_100a) implementation of the one
 * data-parallel hot path of javidsegura/gpu-montecarlo-risk: correlated multi-asset path
 * simulation  R = mu + A z,  S = #{ R_i < -x },  count = #{ trials : S >= k }.
 *
 * Plain pointers and sizes only (no torch / GSL / C++ types), so it can be bound from C
 * (the get_cuda_model plugin, src/06-B200/monte_carlo_b200.c), ctypes, cgo, JNI ...
 * "reference" below = javidsegura/gpu-montecarlo-risk; paths relative to its root.
 *
 * All functions return 0 on success, -1 on failure (the reference's model convention,
 * src/model_interface.h:30 / main_runner.c:248-255); mcr_last_error() describes the failure.
 * There is NO CPU fallback: without a usable CUDA device every compute entry point fails.
 */
#ifndef MCR_B200_H
#define MCR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCR_ABI_VERSION 1

/* ---- transform-matrix conventions (what "L" the reference models really apply) -------- */
#define MCR_TRANSFORM_LOWER 0    /* A = tril(chol(Sigma)): reference CUDA model, src/05-GPU/monte_carlos_cuda.cu:96-108 */
#define MCR_TRANSFORM_GSL_FULL 1 /* A = L + striu(Sigma): what C-serial/OpenMP compute after gsl_linalg_cholesky_decomp1
                                    + a full dgemv, src/02-C-serial/monte_carlo_serial.c:36-37,72 */

/* ---- kernel selection ------------------------------------------------------------------ */
#define MCR_KERNEL_AUTO 0   /* FFMA kernel for N < 36, tensor-core kernels above: one tile kernel up to N = 128, beyond it the
                               single-pass kernel (257..640) or row slabs of 256 assets; MCR_LARGE_N=wide|slab overrides */
#define MCR_KERNEL_FFMA 1   /* fp32 FFMA2 correlate, normals in registers */
#define MCR_KERNEL_TENSOR 2 /* tcgen05 batched Z*A^T (fp16x3 split, fp32 accumulate in TMEM), normals staged in shared memory */
#define MCR_KERNEL_EXACT 3  /* fp64, reference summation order (slow; parity / fix-up kernel) */

#define MCR_FFMA_MAX_N 128

typedef struct mcr_engine mcr_engine; /* opaque; one engine drives one CUDA device */

/* Options of a run.  Zero-initialise, set struct_size = sizeof(mcr_run_options). */
typedef struct mcr_run_options {
    uint32_t struct_size;
    int32_t kernel;        /* MCR_KERNEL_* */
    int32_t no_recheck;    /* 0 (default): per-asset decisions within the fp32 error band of the threshold are
                              recomputed in fp64, so results equal the fp64 reference arithmetic on the same
                              normals bit-for-bit.  1: raw fp32 decisions (benchmark ablation). */
    int32_t want_hist;     /* 1: also fill mcr_result.hist_S (mcr_run only; costs one 4-byte store + load per trial) */
    int32_t want_portfolio;/* 1: also reduce the equal-/custom-weight portfolio return (north-star extension) */
    int32_t sm_reserve;    /* SMs the path kernels leave free (default 0 = none).  The fast kernels are persistent, one CTA per SM with
                              a static share of the tiles, and fill every SM; a collective's kernel launched next to them then waits
                              for an SM and delays the CTA that comes after it - and with it the whole launch.  One process per GPU
                              overlapping each step's all-reduce with the next step's kernels sets 1 (bench.py does). */
    double portfolio_loss_threshold; /* x_p: portfolio tail event is  w.R < -x_p  */
    /* optional per-path output of S (the reference's MonteCarloResult.S_values, src/model_interface.h:12).
     * DEVICE pointer to n_paths int32 (or NULL).  Host callers use mcr_simulate_host instead. */
    int32_t *d_S_values;
} mcr_run_options;

#define MCR_MAX_HIST 1025
#define MCR_PORTFOLIO_BINS 2048

/* Result of a run (or the partial result of one shard of a run). */
typedef struct mcr_result {
    uint64_t n_paths;      /* trials simulated (M of this shard) */
    uint64_t count;        /* #{ trials : S >= k }            (reference: MonteCarloResult.count) */
    uint64_t sum_S;        /* sum of S over trials */
    uint64_t n_rechecked;  /* trials sent through the fp64 fix-up kernel */
    double P_hat, std_error, ci_lower, ci_upper; /* reference formulas, serial.c:152-158 */
    float device_ms;       /* CUDA-event time of the kernel set of this run */
    int32_t kernel_used;   /* MCR_KERNEL_* actually launched */
    int32_t padded_N;      /* asset count the kernel was instantiated for (>= N) */
    int32_t launches;      /* kernels launched by this run */
    /* extension outputs (only when requested) */
    uint64_t hist_S[MCR_MAX_HIST];        /* hist_S[s] = #{ trials : S == s }, s in [0, N] */
    uint64_t portfolio_hits;              /* #{ trials : w.R < -x_p } */
    double portfolio_sum, portfolio_sumsq;/* sum and sum of squares of w.R */
    double portfolio_tail_loss_sum;       /* sum of -(w.R) over tail trials (ES numerator) */
    double portfolio_hist_lo, portfolio_hist_hi; /* histogram range of w.R */
    uint64_t portfolio_hist[MCR_PORTFOLIO_BINS];
    double portfolio_hist_sum[MCR_PORTFOLIO_BINS]; /* per-bin sum of w.R (for ES inside a bin) */
} mcr_result;

/* ---- library ---------------------------------------------------------------------------- */
int mcr_abi_version(void);
/* sizeof of the ABI structs as this library was compiled: 0 = mcr_run_options, 1 = mcr_result, 2 = mcr_host_request */
int mcr_abi_sizeof(int which);
const char *mcr_last_error(void);       /* thread-local, never NULL */
int mcr_device_count(int *n_devices);   /* -1 (not 0 devices) when no CUDA driver/device is usable */

/* ---- host utilities (replace the per-model Cholesky in the reference, serial.c:36-37 /
 *      omp.c:47-48 / cuda.cu:84-108) ------------------------------------------------------- */
/* Sigma: N x N row-major fp64 with leading dimension lds (gsl_matrix.tda).  A_out: N x N dense row-major.
 * mode = MCR_TRANSFORM_*.  Fails (-1) when Sigma is not positive definite. */
int mcr_build_transform(const double *Sigma, int N, int lds, int mode, double *A_out);

/* Pinned (page-locked) host staging for the loader that replaces src/utilities/load_binary.c:13-36,56-87 and for S_values
 * buffers: uploads and the streamed S_values read-back then run at full PCIe / C2C rate without a bounce copy.
 * NULL when no CUDA device is usable or the allocation fails. */
void *mcr_host_alloc_pinned(size_t bytes);
void mcr_host_free_pinned(void *p);

/* ---- engine life cycle ------------------------------------------------------------------ */
int mcr_create(int device, mcr_engine **out);
int mcr_destroy(mcr_engine *e);

/* Upload a model: mu (N fp64) and the dense transform A (N x N row-major fp64).
 * a_is_lower != 0 promises the strict upper triangle of A is zero (triangular kernels skip it). */
int mcr_upload_model(mcr_engine *e, int N, const double *mu, const double *A, int a_is_lower);
/* Optional portfolio weights for the extension outputs (default 1/N each). */
int mcr_set_portfolio_weights(mcr_engine *e, const double *w, int N);

/* ---- statistics and factorisation on the device -------------
 * mcr_device_moments: mean vector and sample covariance (ddof = 1) of `returns` (HOST, T x N row-major fp64): what the
 * reference's offline pipeline computes with pandas (preprocessing/preprocessing_stats.py:18,21) before save_binary.py writes
 * mu.bin / sigma.bin.  mu_out: N, Sigma_out: N x N row-major (HOST).
 * mcr_device_transform: mcr_build_transform on the device - same arguments, same summation order, BITWISE the same A.  */
int mcr_device_moments(mcr_engine *e, const double *returns, int T, int N, double *mu_out, double *Sigma_out);
int mcr_device_transform(mcr_engine *e, const double *Sigma, int N, int lds, int mode, double *A_out);

/* ---- the hot path ----------------------------------------------------------------------- */
/* Simulate trials [path_offset, path_offset + M) of the stream keyed by seed.  Normals come from
 * Philox4x32-10 with key = seed and counter = (path index, asset block), so any partition of the
 * path range over calls / GPUs gives bit-identical totals.  Synchronous; fills *out. */
int mcr_run(mcr_engine *e, uint64_t M, uint64_t path_offset, uint64_t seed, double x, int k,
            const mcr_run_options *opt, mcr_result *out);

/* Asynchronous form for callers that own the stream and combine shards themselves
 * (one process per GPU + NCCL): enqueue the kernel set on `cuda_stream` (a cudaStream_t; NULL =
 * the engine's own stream) and leave the partials in d_partials (DEVICE pointer to
 * MCR_PARTIAL_WORDS uint64: [0]=n_paths [1]=count [2]=sum_S [3]=n_rechecked [4]=status, rest reserved).
 * The words are plain sums, so a sum-allreduce of the buffer across ranks is the cross-GPU step. */
#define MCR_PARTIAL_WORDS 8
int mcr_run_async(mcr_engine *e, uint64_t M, uint64_t path_offset, uint64_t seed, double x, int k,
                  const mcr_run_options *opt, uint64_t *d_partials, void *cuda_stream);
/* Device time of the dominant (path) kernel of the engine's most recent run: average CUDA-event duration per
 * launch, and the number of launches timed.  Waits for that run to finish. */
int mcr_last_main_kernel_ms(mcr_engine *e, float *avg_ms_per_launch, int *n_launches);
/* fp16 MMA flops per trial the tensor pipe executed in the engine's most recent run (3 passes of the fp16 hi/lo split over
 * the padded tiles actually issued; 0 for the FFMA and fp64 kernels): the "executed" figure next to the algorithmic
 * F(N) = N^2 + 3N of the roofline in bench.py. */
int mcr_last_mma_flops_per_trial(mcr_engine *e, double *flops);
/* Turn (possibly all-reduced) partial words copied to the host into a result with the reference's statistics. */
int mcr_result_from_partials(const uint64_t *h_partials, mcr_result *out);

/* Extension outputs (want_hist / want_portfolio) of a shard, for callers that combine shards themselves: same as
 * mcr_run_async, and the extension partials of the run are left in two DEVICE buffers of plain sums, so the
 * cross-GPU step is one more sum-allreduce per buffer (uint64 and fp64):
 *   d_ext_counts[MCR_EXT_COUNT_WORDS] uint64: hist_S[0..MCR_MAX_HIST) | portfolio_hits | portfolio_hist[0..BINS)
 *   d_ext_sums[MCR_EXT_SUM_WORDS]     fp64  : sum w.R, sum (w.R)^2, tail-loss sum | portfolio_hist_sum[0..BINS)
 * Parts that were not requested are zero.  The portfolio histogram range depends on the model and the weights only
 * (w.mu +- 8 sigma_p), so every shard bins identically; mcr_portfolio_hist_range returns it after the run was enqueued. */
#define MCR_EXT_COUNT_WORDS (MCR_MAX_HIST + 1 + MCR_PORTFOLIO_BINS)
#define MCR_EXT_SUM_WORDS (3 + MCR_PORTFOLIO_BINS)
int mcr_run_async_ext(mcr_engine *e, uint64_t M, uint64_t path_offset, uint64_t seed, double x, int k,
                      const mcr_run_options *opt, uint64_t *d_partials, uint64_t *d_ext_counts, double *d_ext_sums,
                      void *cuda_stream);
int mcr_portfolio_hist_range(mcr_engine *e, double *lo, double *hi);
/* mcr_result_from_partials plus the extension fields, from (possibly all-reduced) host copies of the three buffers. */
int mcr_result_from_partials_ext(const uint64_t *h_partials, const uint64_t *h_ext_counts, const double *h_ext_sums,
                                 double portfolio_hist_lo, double portfolio_hist_hi, mcr_result *out);

/* Parity entry point (the model interface has no "feed normals" hook):
 * same correlate/threshold/reduce code, normals supplied by the caller instead of Philox.
 * Z: HOST pointer, M x N row-major fp64 (the layout serial.c:67-69 fills, one trial per row).
 * Kernel choice under MCR_KERNEL_AUTO: the FFMA2 kernel for N < 64 and for odd N <= 128, the streamed tcgen05 kernel otherwise
 * (even N: it reads the matrix through a 2-D TMA tensor map, which needs 16-byte aligned rows; any other matrix is read
 * element by element, slower but identical).  Portfolio outputs are not available in this mode. */
int mcr_run_with_normals(mcr_engine *e, uint64_t M, const double *Z, double x, int k,
                         const mcr_run_options *opt, int32_t *S_values_host /* M int32 or NULL */,
                         mcr_result *out);
/* Same with DEVICE-resident normals (bench: HBM-resident input); d_S_values via opt. */
int mcr_run_with_normals_device(mcr_engine *e, uint64_t M, const double *d_Z, double x, int k,
                                const mcr_run_options *opt, mcr_result *out);

/* Debug/parity: materialise what the in-register generator produces for paths
 * [path_offset, path_offset+M): normals (M x N fp32, HOST) and/or raw Philox words
 * (M x 4*ceil(N/4) uint32, HOST).  Either pointer may be NULL. */
int mcr_dump_normals(mcr_engine *e, uint64_t M, uint64_t path_offset, uint64_t seed,
                     float *Z_host, uint32_t *U_host);

/* ---- one-shot host entry point (what the get_cuda_model plugin calls) ------------------- */
typedef struct mcr_host_request {
    uint32_t struct_size;
    int32_t N, k;
    uint64_t M;
    uint64_t seed;
    double x;
    const double *mu;      /* N */
    const double *A;       /* N x N row-major dense transform */
    int32_t a_is_lower;
    int32_t n_gpus;        /* 0/1 = one device; >1 = shard the path range over that many devices of this
                              process and combine with one NCCL all-reduce */
    int32_t first_device;
    int32_t want_S_values; /* fill S_values_out */
    int32_t reupload;      /* 1: copy the model host->device even if this process already uploaded an identical one */
    int32_t reserved1;
    int32_t *S_values_out; /* HOST, M int32 (may be NULL) */
    mcr_run_options opt;
} mcr_host_request;
int mcr_simulate_host(const mcr_host_request *req, mcr_result *out);

/* Extension (north-star A7, no reference counterpart): value-at-risk and expected shortfall of the portfolio
 * loss -w.R at confidence alpha (e.g. 0.99), from the merged histogram of a want_portfolio run.  VaR is exact
 * to one bin width ((hist_hi - hist_lo) / MCR_PORTFOLIO_BINS = sigma_p / 128); ES uses the exact per-bin sums. */
int mcr_portfolio_var_es(const mcr_result *r, double alpha, double *var_out, double *es_out);

/* FP32-pipe calibration used by bench.py for the roofline denominator: runs a pure FFMA2 kernel
 * (no memory traffic) for about `ms_target` ms and reports achieved fp32 TFLOP/s. */
int mcr_measure_fp32_peak(mcr_engine *e, float ms_target, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* MCR_B200_H */
