// C ABI of libmcr_b200.so (declared in include/mcr_b200.h).
//
// Host-side orchestration only: model upload (fp64 -> packed fp32 constant layout + fp64
// column-major copy for the fix-up), kernel selection, chunked launches
//   [ffma_path_kernel | exact_kernel] -> exact_kernel(fix-up list) -> finalize_kernel
// on one stream, and the reference's statistics (serial.c:152-158).  No CPU compute fallback.
#include "../../include/mcr_b200.h"
#include "mcr_internal.h"
#include "mcr_engine.h"
#include "mcr_rng.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace mcr {

// ---- error reporting ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return -1;
}

// ---- variant registry ----------------------------------------------------------------------------
VariantRegistry& registry() {
    static VariantRegistry r;
    return r;
}
void register_variant(const FfmaVariant& v) {
    VariantRegistry& r = registry();
    if (r.n < MAX_VARIANTS) r.v[r.n++] = v;
}
static void ensure_registered() {
    static std::once_flag once;
    std::call_once(once, [] {
        register_group0(); register_group1(); register_group2(); register_group3();
        register_group4(); register_group5(); register_group6(); register_group7();
    });
}
// smallest instantiation that can run an N-asset model; a lower-triangular model may also use a
// dense instantiation (zeros are stored), never the other way round.
static int find_variant(int N, bool lower, bool normals) {
    const VariantRegistry& r = registry();
    int best = -1;
    for (int i = 0; i < r.n; ++i) {
        const FfmaVariant& v = r.v[i];
        if (v.normals != normals || v.NT < N) continue;
        if (!lower && v.lower) continue;
        if (best < 0 || v.NT < r.v[best].NT || (v.NT == r.v[best].NT && v.lower && !r.v[best].lower)) best = i;
    }
    return best;
}

// which (model generation, variant) currently sits in each constant storage of each device
struct ConstState { uint64_t gen = 0; int variant = -1; uint64_t c_gen = 0; int c_variant = -1; double c_x = 0; };
static std::mutex g_const_mu;
static std::map<std::pair<int, const void*>, ConstState> g_const_state;
static std::atomic<uint64_t> g_next_gen{1};


}  // namespace mcr

using namespace mcr;


static int set_device(const mcr_engine* e) {
    MCR_CUDA(cudaSetDevice(e->device));
    return 0;
}

extern "C" {

int mcr_abi_version(void) { return MCR_ABI_VERSION; }
int mcr_abi_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(mcr_run_options);
        case 1: return (int)sizeof(mcr_result);
        case 2: return (int)sizeof(mcr_host_request);
        default: return -1;
    }
}
const char* mcr_last_error(void) { return g_err; }

int mcr_device_count(int* n) {
    if (!n) return fail("mcr_device_count: null argument");
    int c = 0;
    cudaError_t err = cudaGetDeviceCount(&c);
    if (err != cudaSuccess || c <= 0) {
        *n = 0;
        return fail("no usable CUDA device: %s", err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
    }
    *n = c;
    return 0;
}

// ---- host Cholesky / transform (fp64) ------------------------------------------------------------
int mcr_build_transform(const double* Sigma, int N, int lds, int mode, double* A) {
    if (!Sigma || !A || N <= 0 || lds < N) return fail("mcr_build_transform: bad arguments");
    if (mode != MCR_TRANSFORM_LOWER && mode != MCR_TRANSFORM_GSL_FULL) return fail("mcr_build_transform: unknown mode %d", mode);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) A[(size_t)i * N + j] = Sigma[(size_t)i * lds + j];
    // left-looking Cholesky on the lower triangle, every inner product accumulated left to right (one rounding per
    // multiply and per add: bitwise the unblocked reference algorithm); eight rows at a time so that the eight
    // independent chains hide the add latency (N = 1024: 180 ms -> 30 ms on one host core)
    for (int j = 0; j < N; ++j) {
        const double* aj = A + (size_t)j * N;
        int i = j;
        for (; i + 8 <= N; i += 8) {
            const double *a0 = A + (size_t)i * N, *a1 = a0 + N, *a2 = a1 + N, *a3 = a2 + N, *a4 = a3 + N, *a5 = a4 + N, *a6 = a5 + N, *a7 = a6 + N;
            double t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7 = 0;
            for (int k = 0; k < j; ++k) {
                const double v = aj[k];
                t0 += v * a0[k]; t1 += v * a1[k]; t2 += v * a2[k]; t3 += v * a3[k];
                t4 += v * a4[k]; t5 += v * a5[k]; t6 += v * a6[k]; t7 += v * a7[k];
            }
            const double t[8] = {t0, t1, t2, t3, t4, t5, t6, t7};
            for (int r = 0; r < 8; ++r) A[(size_t)(i + r) * N + j] -= t[r];
        }
        for (; i < N; ++i) {
            double t = 0.0;
            for (int k = 0; k < j; ++k) t += aj[k] * A[(size_t)i * N + k];
            A[(size_t)i * N + j] -= t;
        }
        const double ajj = A[(size_t)j * N + j];
        if (!(ajj > 0.0) || !std::isfinite(ajj)) return fail("mcr_build_transform: Sigma is not positive definite (pivot %d = %g)", j, ajj);
        const double inv = 1.0 / std::sqrt(ajj);
        for (int i = j; i < N; ++i) A[(size_t)i * N + j] *= inv;
    }
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j)
            A[(size_t)i * N + j] = (mode == MCR_TRANSFORM_GSL_FULL) ? Sigma[(size_t)j * lds + i] : 0.0;
    return 0;
}

void* mcr_host_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        fail("mcr_host_alloc_pinned: cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void mcr_host_free_pinned(void* p) { if (p) cudaFreeHost(p); }

// ---- engine --------------------------------------------------------------------------------------
int mcr_create(int device, mcr_engine** out) {
    if (!out) return fail("mcr_create: null out");
    *out = nullptr;
    int n = 0;
    if (mcr_device_count(&n) != 0) return -1;
    if (device < 0 || device >= n) return fail("mcr_create: device %d out of range (have %d)", device, n);
    ensure_registered();
    MCR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MCR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail("mcr_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    mcr_engine* e = new mcr_engine();
    e->device = device;
    e->num_sms = prop.multiProcessorCount;
    cudaError_t err = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
    ok(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    ok(cudaEventCreate(&e->ev0));
    ok(cudaEventCreate(&e->ev1));
    ok(cudaEventCreateWithFlags(&e->ev_run_done, cudaEventDisableTiming));
    for (int i = 0; i < 2 * MAX_TIMED_CHUNKS; ++i) ok(cudaEventCreate(&e->evk[i]));
    ok(cudaMalloc(&e->d_flag_list, sizeof(uint32_t) * FLAG_CAP));
    ok(cudaMalloc(&e->d_flag_count, sizeof(uint32_t)));
    ok(cudaMalloc(&e->d_pair_list, sizeof(unsigned long long) * PAIR_CAP));
    ok(cudaMalloc(&e->d_pair_count, sizeof(uint32_t)));
    ok(cudaMalloc(&e->d_accum, sizeof(unsigned long long) * 2));
    ok(cudaMalloc(&e->d_block_partials, sizeof(uint64_t) * 2 * MAX_GRID));
    ok(cudaMalloc(&e->d_partials, sizeof(uint64_t) * MCR_PARTIAL_WORDS));
    ok(cudaMalloc(&e->d_hist, sizeof(unsigned long long) * MCR_MAX_HIST));
    ok(cudaMalloc(&e->d_port_acc, sizeof(double) * 4));
    ok(cudaMalloc(&e->d_port_hist, sizeof(unsigned long long) * MCR_PORTFOLIO_BINS));
    ok(cudaMalloc(&e->d_port_hist_sum, sizeof(double) * MCR_PORTFOLIO_BINS));
    ok(cudaMallocHost(&e->h_partials, sizeof(uint64_t) * MCR_PARTIAL_WORDS));
    if (err != cudaSuccess) {
        mcr_destroy(e);
        return fail("mcr_create: allocation failed: %s", cudaGetErrorString(err));
    }
    *out = e;
    return 0;
}

int mcr_destroy(mcr_engine* e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->d_mu); cudaFree(e->d_Acm); cudaFree(e->d_tc_image); cudaFree(e->d_slab_image); cudaFree(e->d_S_acc);
    cudaFree(e->d_pair_list); cudaFree(e->d_pair_count); cudaFree(e->d_Arm); cudaFree(e->d_wide_image); cudaFree(e->d_wide_c); cudaFree(e->d_stream_image);
    cudaFree(e->d_flag_list); cudaFree(e->d_flag_count); cudaFree(e->d_accum);
    cudaFree(e->d_block_partials); cudaFree(e->d_partials);
    cudaFree(e->d_port_acc); cudaFree(e->d_port_hist); cudaFree(e->d_port_hist_sum);
    cudaFree(e->d_hist); cudaFree(e->d_S_tmp); cudaFree(e->d_ext_counts); cudaFree(e->d_ext_sums);
    if (e->h_partials) cudaFreeHost(e->h_partials);
    cudaFree(e->d_S_part);
    for (int b = 0; b < 2; ++b) {
        cudaFree(e->d_S_compact[b]);
        if (e->h_S_compact[b]) cudaFreeHost(e->h_S_compact[b]);
        if (e->ev_compact[b]) cudaEventDestroy(e->ev_compact[b]);
        if (e->ev_copied[b]) cudaEventDestroy(e->ev_copied[b]);
    }
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->ev_run_done) cudaEventDestroy(e->ev_run_done);
    for (int i = 0; i < 2 * MAX_TIMED_CHUNKS; ++i) if (e->evk[i]) cudaEventDestroy(e->evk[i]);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return 0;
}

int mcr_upload_model(mcr_engine* e, int N, const double* mu, const double* A, int a_is_lower) {
    if (!e || !mu || !A) return fail("mcr_upload_model: null argument");
    if (N <= 0 || N >= MCR_MAX_HIST) return fail("mcr_upload_model: N=%d out of range [1, %d]", N, MCR_MAX_HIST - 1);
    if (set_device(e)) return -1;
    for (int i = 0; i < N; ++i) {
        if (!std::isfinite(mu[i])) return fail("mcr_upload_model: mu[%d] is not finite", i);
        for (int j = 0; j < N; ++j) {
            const double a = A[(size_t)i * N + j];
            if (!std::isfinite(a)) return fail("mcr_upload_model: A[%d][%d] is not finite", i, j);
            if (a_is_lower && j > i && a != 0.0) return fail("mcr_upload_model: a_is_lower set but A[%d][%d]=%g", i, j, a);
        }
    }
    MCR_CUDA(cudaStreamSynchronize(e->stream));
    if (e->run_pending) { MCR_CUDA(cudaEventSynchronize(e->ev_run_done)); e->run_pending = false; }   // a run on a caller's stream
    // device copies first (fp64 for the exact / fix-up kernels: A column-major and row-major); the model is published
    // - N, generation, host copies - only once they are in place, so a failed upload leaves no half-uploaded model behind
    e->N = 0;
    if (N != e->N_alloc) {
        cudaFree(e->d_mu); cudaFree(e->d_Acm); cudaFree(e->d_Arm);
        e->d_mu = nullptr; e->d_Acm = nullptr; e->d_Arm = nullptr; e->N_alloc = 0;
        MCR_CUDA(cudaMalloc(&e->d_mu, sizeof(double) * N));
        MCR_CUDA(cudaMalloc(&e->d_Acm, sizeof(double) * (size_t)N * N));
        MCR_CUDA(cudaMalloc(&e->d_Arm, sizeof(double) * (size_t)N * N));
        e->N_alloc = N;
    }
    std::vector<double> Acm((size_t)N * N);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) Acm[(size_t)j * N + i] = A[(size_t)i * N + j];
    MCR_CUDA(cudaMemcpy(e->d_mu, mu, sizeof(double) * N, cudaMemcpyHostToDevice));
    MCR_CUDA(cudaMemcpy(e->d_Acm, Acm.data(), sizeof(double) * (size_t)N * N, cudaMemcpyHostToDevice));
    MCR_CUDA(cudaMemcpy(e->d_Arm, A, sizeof(double) * (size_t)N * N, cudaMemcpyHostToDevice));   // row-major, per-decision fix-up
    e->lower = a_is_lower != 0;
    e->mu.assign(mu, mu + N);
    e->A.assign(A, A + (size_t)N * N);
    e->w.assign(N, 1.0 / N);
    e->packed.assign(registry().n, std::vector<float>());
    e->gen = g_next_gen.fetch_add(1);
    e->N = N;
    // guard-band ingredient: gamma_i * sum_j |A_ij| with gamma_i = (terms_i + 3) * 2^-24
    const double u = std::ldexp(1.0, -24);
    e->gamma_rowabs_max = 0;
    for (int i = 0; i < N; ++i) {
        double s = 0;
        for (int j = 0; j < N; ++j) s += std::fabs(A[(size_t)i * N + j]);
        const int terms = e->lower ? i + 1 : N;
        e->gamma_rowabs_max = std::max(e->gamma_rowabs_max, (terms + 3) * u * s);
    }
    return 0;
}

int mcr_set_portfolio_weights(mcr_engine* e, const double* w, int N) {
    if (!e || !w) return fail("mcr_set_portfolio_weights: null argument");
    if (N != e->N) return fail("mcr_set_portfolio_weights: N=%d but the uploaded model has N=%d", N, e->N);
    e->w.assign(w, w + N);
    return 0;
}

}  // extern "C"

// ---- launch plumbing -----------------------------------------------------------------------------
namespace mcr {

static float round_up_to_float(double v) {
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, INFINITY);
    return f;
}

struct RunPlan {
    int kernel;       // MCR_KERNEL_*
    int variant;      // FFMA variant index or -1
    int NT;
    bool recheck;
    bool portfolio;   // accumulator row N carries the portfolio return (tensor kernel only)
    bool slab;        // N > 128: tensor kernel over row slabs of 256 assets (MCR_LARGE_N=slab)
    bool wide;        // N > 128: single-pass tensor kernel of round 1 (MCR_LARGE_N=wide)
    bool stream;      // streamed tensor kernel: normals in TMEM, passes over row ranges (N > 128 by default; supplied normals)
    int rows;         // accumulator rows needed: N (+1 with the portfolio row)
};

static int auto_tc_min_n() {
    static const int v = [] { const char* e = getenv("MCR_AUTO_TC_MIN_N"); return e && *e ? atoi(e) : 36; }();
    return v;
}

static int make_plan(const mcr_engine* e, const mcr_run_options* opt, bool normals, RunPlan* plan) {
    const int want = opt ? opt->kernel : MCR_KERNEL_AUTO;
    *plan = RunPlan{};
    plan->recheck = !(opt && opt->no_recheck);
    plan->variant = -1;
    plan->NT = e->N;
    plan->rows = e->N;
    if (want == MCR_KERNEL_EXACT) { plan->kernel = MCR_KERNEL_EXACT; return 0; }
    plan->portfolio = opt && opt->want_portfolio;
    plan->slab = false;
    plan->rows = e->N + (plan->portfolio ? 1 : 0);
    if (plan->portfolio && normals) return fail("portfolio outputs are not available with supplied normals");
    if (plan->rows > tc_max_n() && (want == MCR_KERNEL_TENSOR || want == MCR_KERNEL_AUTO) && stream_supported(e->N, plan->portfolio ? 1 : 0)) {
        // more accumulator rows than the tile kernel holds (N > 128, or N = 128 plus the portfolio row), generated or supplied
        // normals: the streamed kernel (mcr_stream.cu).  The two round-1 kernels stay selectable for A/B
        // runs: MCR_LARGE_N=slab|wide (generated normals, no portfolio outputs).
        static const int forced = [] { const char* v = getenv("MCR_LARGE_N"); return !v ? 0 : !strcmp(v, "slab") ? 1 : !strcmp(v, "wide") ? 2 : 0; }();
        plan->kernel = MCR_KERNEL_TENSOR;
        const bool r1_ok = !normals && !plan->portfolio && e->N > tc_max_n();
        if (r1_ok && forced == 2 && wide_supported(e->N)) { plan->wide = true; plan->NT = 128 * wide_slabs(e->N); }
        else if (r1_ok && forced == 1) { plan->slab = true; plan->NT = slab_rows() * slab_count(e->N); }
        else { plan->stream = true; plan->NT = (plan->rows + 15) & ~15; }
        return 0;
    }
    const bool tc_ok = tc_supported(plan->rows);   // (supplied normals: the tile kernel's supplied-normals instantiation, on request)
    if (plan->portfolio) {
        if (want != MCR_KERNEL_AUTO && want != MCR_KERNEL_TENSOR) return fail("portfolio outputs are produced by the tensor-core kernel only");
        if (!tc_ok) return fail("portfolio outputs: N=%d is outside the tensor kernels' range", e->N);
        plan->kernel = MCR_KERNEL_TENSOR; plan->NT = tc_padded_n(plan->rows);
        return 0;
    }
    // Supplied normals, N <= 128.  The streamed kernel reads the caller's matrix through TMA boxes (even N: rows 16-byte aligned) at
    // 2.0 - 3.0 TB/s for N = 50 ... 128, the tile kernel's supplied-normals instantiation (element loads) at 0.8 - 1.6, the FFMA kernel
    // at 4.3 (N <= 32), 3.3 (N = 50), 2.1 (N = 64), 1.4 (N = 100), 0.9 (N = 128) (profiles/r02_supplied_normals_bandwidth.log):
    // on request (TENSOR) even N >= 36 takes the streamed kernel, AUTO takes it from N = 64.  MCR_SUPPLIED_TENSOR=tile keeps the tile kernel.
    static const bool supplied_tile = [] { const char* v = getenv("MCR_SUPPLIED_TENSOR"); return v && !strcmp(v, "tile"); }();
    if (normals && !supplied_tile && (e->N & 1) == 0 && stream_supported(e->N, 0) &&
        ((want == MCR_KERNEL_TENSOR && e->N >= 36) || (want == MCR_KERNEL_AUTO && e->N >= 64))) {
        plan->kernel = MCR_KERNEL_TENSOR; plan->stream = true; plan->NT = (plan->rows + 15) & ~15;
        return 0;
    }
    if (want == MCR_KERNEL_TENSOR) {
        if (!tc_ok) return fail("tensor-core kernel covers N <= %d in this build (N=%d)", tc_max_n(), e->N);
        plan->kernel = MCR_KERNEL_TENSOR; plan->NT = normals ? tc_supplied_padded_n() : tc_padded_n(plan->rows);
        return 0;
    }
    const int v = find_variant(e->N, e->lower, normals);
    // AUTO: the FFMA kernel while the transform (roughly) fits the SM's first-level constant cache, the
    // tensor-core kernel above that (the FFMA kernel then stalls on constant-cache misses).  Measured
    // crossover on B200: N = 32 FFMA 2.3e10 vs tensor 2.1e10 sims/s, N = 40 1.6e10 vs 1.7e10
    // (profiles/r01_sweep_ffma_vs_tensor.log)
    if (want == MCR_KERNEL_AUTO && tc_ok && !normals && (e->N >= auto_tc_min_n() || v < 0)) {
        plan->kernel = MCR_KERNEL_TENSOR; plan->NT = tc_padded_n(plan->rows);
        return 0;
    }
    if (v >= 0) {
        plan->kernel = MCR_KERNEL_FFMA; plan->variant = v; plan->NT = registry().v[v].NT;
        return 0;
    }
    if (want == MCR_KERNEL_FFMA)
        return fail("no FFMA instantiation for N=%d (%s transform); FFMA path covers N <= %d lower / N <= 100 dense",
                    e->N, e->lower ? "lower" : "dense", MCR_FFMA_MAX_N);
    plan->kernel = MCR_KERNEL_EXACT;  // AUTO beyond the fast kernels' range: fp64 kernel (still GPU)
    return 0;
}

// enqueue one full run on stream s; partials accumulate into d_out (zeroed here)
int enqueue_run(mcr_engine* e, uint64_t M, uint64_t path_offset, uint64_t seed, double x, int k,
                       const mcr_run_options* opt, const double* dZ, uint64_t* d_out, cudaStream_t s) {
    if (e->N <= 0) return fail("no model uploaded (call mcr_upload_model first)");
    if (!std::isfinite(x)) return fail("x is not finite");
    if (set_device(e)) return -1;
    RunPlan plan;
    if (make_plan(e, opt, dZ != nullptr, &plan)) return -1;
    const int N = e->N;
    // S >= k with k <= 0 holds for every trial (the reference compares signed ints, serial.c:137-141); the fast kernels compare
    // unsigned, so clamp once here and every kernel of the run agrees with the reference
    if (k < 0) k = 0;
    // One run in flight per engine: every run shares the engine's scratch (fix-up lists, accumulators, block partials,
    // operand images, the FFMA constant bank).  A run enqueued on another stream waits for the previous one.
    if (e->run_pending) MCR_CUDA(cudaStreamWaitEvent(s, e->ev_run_done, 0));
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    int32_t* d_S = opt ? opt->d_S_values : nullptr;
    int launches = 0;
    double mma_flops = 0;   // fp16 MMA flops the tensor pipe executes per trial (0 on the FFMA / fp64 kernels)
    // histogram of S: needs the per-trial S of every chunk; use the caller's buffer or a per-chunk scratch one
    const bool want_hist = opt && opt->want_hist;
    bool S_is_scratch = false;
    // SMs the path kernels may fill: all of them, less the ones the caller keeps free for a collective that runs concurrently
    const int sms = std::max(2, e->num_sms - ((opt && opt->sm_reserve > 0) ? std::min(opt->sm_reserve, e->num_sms / 2) : 0));
    const bool accumulate = e->accumulate;   // a later part of a run that mcr_simulate_host enqueues in pieces: add to the outputs
    if (want_hist) {
        if (!accumulate) MCR_CUDA(cudaMemsetAsync(e->d_hist, 0, sizeof(unsigned long long) * MCR_MAX_HIST, s));
        if (!d_S) {
            const uint64_t need = std::min<uint64_t>(M, CHUNK_PATHS);
            if (e->S_tmp_cap < need) {
                cudaFree(e->d_S_tmp); e->d_S_tmp = nullptr; e->S_tmp_cap = 0;
                MCR_CUDA(cudaMalloc(&e->d_S_tmp, sizeof(int32_t) * std::max<uint64_t>(need, 1)));
                e->S_tmp_cap = need;
            }
            d_S = e->d_S_tmp;
            S_is_scratch = true;
        }
    }
    e->last_hist = want_hist;
    if (!accumulate) {
        e->n_timed = 0;
        MCR_CUDA(cudaMemsetAsync(d_out, 0, sizeof(uint64_t) * MCR_PARTIAL_WORDS, s));
    }

    const FfmaVariant* var = plan.variant >= 0 ? &registry().v[plan.variant] : nullptr;
    int grid_full = 0;
    FfmaParams fp{};
    if (var) {
        // thresholds c_i = mu_i + x (padded rows can never hit) and the guard band
        std::vector<float> c(plan.NT, 1.0f);
        double cabs_gamma_max = 0;
        const double u = std::ldexp(1.0, -24);
        for (int i = 0; i < N; ++i) {
            const double ci = e->mu[i] + x;
            c[i] = (float)ci;
            const int terms = e->lower ? i + 1 : N;
            cabs_gamma_max = std::max(cabs_gamma_max, (terms + 3) * u * std::fabs(ci));
        }
        const double safety = 2.0;
        if (plan.recheck) {
            fp.delta = round_up_to_float(safety * (cabs_gamma_max + (double)BOX_MULLER_ZMAX * e->gamma_rowabs_max) + 1e-30);
            fp.delta_rel = round_up_to_float(safety * e->gamma_rowabs_max);
            fp.delta_abs = round_up_to_float(safety * cabs_gamma_max + 1e-30);
        } else {
            fp.delta = 0.f; fp.delta_rel = 0.f; fp.delta_abs = 0.f;
        }
        // constant storage: transform (only when this device's copy is stale) and thresholds
        {
            std::lock_guard<std::mutex> lk(g_const_mu);
            ConstState& st = g_const_state[{e->device, var->symbol_tag}];
            if (st.gen != e->gen || st.variant != plan.variant) {
                std::vector<float>& packed = e->packed[plan.variant];
                if (packed.empty()) {
                    packed.resize(2 * (size_t)var->n_pairs);
                    var->pack(e->A.data(), N, packed.data());
                }
                // a previous launch on another stream may still read the old contents
                MCR_CUDA(cudaDeviceSynchronize());
                MCR_CUDA(var->upload_A(packed.data(), s));
                st.gen = e->gen; st.variant = plan.variant;
            }
            if (st.c_gen != e->gen || st.c_variant != plan.variant || st.c_x != x) {
                // pageable source: the runtime stages it before returning, so `c` may go out of scope
                MCR_CUDA(var->upload_c(c.data(), s));
                st.c_gen = e->gen; st.c_variant = plan.variant; st.c_x = x;
            }
        }
        int bps = 0;
        MCR_CUDA(var->occupancy(128, &bps));
        if (bps < 1) return fail("FFMA kernel NT=%d cannot be resident", plan.NT);
        grid_full = std::min(sms * bps, MAX_GRID);
        fp.key0 = key0; fp.key1 = key1; fp.k = k; fp.N = N;
        fp.flag_list = e->d_flag_list; fp.flag_count = e->d_flag_count; fp.flag_cap = FLAG_CAP;
        fp.block_partials = e->d_block_partials;
    }

    // portfolio extension: the extra image row (w^T A, then w.mu), the histogram range (w.mu +- 8 sigma_p: a function of the model
    // and the weights only, so every shard and GPU bins identically), zeroed accumulators
    auto prepare_portfolio = [&](std::vector<double>& port_row, double& hist_lo, double& hist_width) -> int {
        port_row.assign(N + 1, 0.0);
        double var = 0;
        for (int j = 0; j < N; ++j) {
            double vj = 0;
            for (int i = 0; i < N; ++i) vj += e->w[i] * e->A[(size_t)i * N + j];
            port_row[j] = vj;
            var += vj * vj;
        }
        for (int i = 0; i < N; ++i) port_row[N] += e->w[i] * e->mu[i];
        const double sd = std::sqrt(var);
        hist_lo = port_row[N] - 8.0 * sd;
        hist_width = (16.0 * sd > 0 ? 16.0 * sd : 1.0) / MCR_PORTFOLIO_BINS;
        e->port_hist_lo = hist_lo; e->port_hist_hi = hist_lo + hist_width * MCR_PORTFOLIO_BINS;
        if (!accumulate) {
            MCR_CUDA(cudaMemsetAsync(e->d_port_acc, 0, sizeof(double) * 4, s));
            MCR_CUDA(cudaMemsetAsync(e->d_port_hist, 0, sizeof(unsigned long long) * MCR_PORTFOLIO_BINS, s));
            MCR_CUDA(cudaMemsetAsync(e->d_port_hist_sum, 0, sizeof(double) * MCR_PORTFOLIO_BINS, s));
        }
        return 0;
    };

    TcLaunch tl{};
    int tc_grid_full = 0;
    if (plan.kernel == MCR_KERNEL_TENSOR && !plan.slab && !plan.wide && !plan.stream) {
        // split, swizzled transform image (depends on x through the constant column): rebuilt only when
        // the model or x changed
        const int rows = plan.rows;
        const bool sup = dZ != nullptr;   // the supplied-normals instantiation: one shape (128 rows, 3 K-blocks), its own image
        const size_t bytes = sup ? tc_supplied_image_bytes() : tc_image_bytes(rows);
        const int nout = sup ? tc_supplied_padded_n() : tc_padded_n(rows);
        if (e->tc_image_cap < bytes) {
            cudaFree(e->d_tc_image); e->d_tc_image = nullptr; e->tc_image_cap = 0;
            MCR_CUDA(cudaMalloc(&e->d_tc_image, bytes));
            e->tc_image_cap = bytes; e->tc_gen = 0;
        }
        std::vector<double> c(N);
        for (int i = 0; i < N; ++i) c[i] = e->mu[i] + x;
        // portfolio row: P = w.R = w.mu + (w^T A) z rides through the same MMAs as one more output row
        std::vector<double> port_row;
        if (plan.portfolio) {
            if (prepare_portfolio(port_row, tl.hist_lo, tl.hist_width)) return -1;
            tl.portfolio = 1;
            tl.port_x = opt->portfolio_loss_threshold;
            tl.port_acc = e->d_port_acc; tl.port_hist = e->d_port_hist; tl.port_hist_sum = e->d_port_hist_sum;
        }
        const uint64_t image_key = e->gen * 4 + (plan.portfolio ? 1 : 0) + (sup ? 2 : 0);
        if (e->tc_gen != image_key || e->tc_x != x || plan.portfolio) {
            e->tc_host_image.resize(bytes);
            e->tc_row_scale.assign(nout, 1.0);
            if (sup) tc_supplied_build(N, e->A.data(), c.data(), e->tc_host_image.data(), e->tc_row_scale.data());
            else tc_build(N, rows, e->A.data(), c.data(), plan.portfolio ? port_row.data() : nullptr, e->tc_host_image.data(), e->tc_row_scale.data());
            MCR_CUDA(cudaStreamSynchronize(s));   // a previous run may still read the old image
            MCR_CUDA(cudaMemcpyAsync(e->d_tc_image, e->tc_host_image.data(), bytes, cudaMemcpyHostToDevice, s));
            e->tc_gen = image_key; e->tc_x = x;
        }
        if (plan.portfolio) tl.port_scale = e->tc_row_scale[nout - 1];
        // Guard band, in accumulator units (accumulator i = s_i (R_i + x), s_i the row's power-of-two scale).  fp16x3: each
        // operand is hi + lo with relative residual 2^-22 (plus fp16-subnormal rounding of tiny z_lo, absolute 2^-25 per
        // normal), the lo*lo product is dropped, and the tensor core accumulates in fp32 over 12*KB MMA instructions: budget
        // 2^-19 relative to |c| + sum_j |a_ij||z_j| (3 * 2^-22 for the split, the rest for accumulation), doubled like the
        // FFMA band (derivation and the assumed accumulation model: DESIGN.md section 3.1).  Calibration on 4M trials per size:
        // counts are still exact with the band at 1/64 of this (profiles/r01_tc_band_calibrate.log).
        static const double band_scale = [] { const char* v = getenv("MCR_TC_BAND_SCALE"); return v && *v ? atof(v) : 1.0; }();
        const double gamma = std::ldexp(1.0, -19);
        double cabs_max = 0, rowabs_max = 0;
        for (int i = 0; i < N; ++i) {
            double srow = 0;
            for (int j = 0; j < N; ++j) srow += std::fabs(e->A[(size_t)i * N + j]);
            rowabs_max = std::max(rowabs_max, srow * e->tc_row_scale[i]);
            cabs_max = std::max(cabs_max, std::fabs(c[i]) * e->tc_row_scale[i]);
        }
        tl.delta_rel = plan.recheck ? band_scale * 2.0 * gamma * rowabs_max : 0.0;
        tl.delta_abs = plan.recheck ? band_scale * 2.0 * gamma * cabs_max + 1e-30 : 0.0;
        int bps = 0;
        MCR_CUDA(sup ? tc_supplied_occupancy(&bps) : tc_occupancy(plan.rows, plan.portfolio, &bps));
        if (bps < 1) return fail("tensor-core kernel for N=%d%s does not fit one SM's shared memory", N, plan.portfolio ? " with the portfolio histogram" : "");
        if (const char* ov = getenv("MCR_TC_CTAS_PER_SM")) { if (atoi(ov) > 0) bps = atoi(ov); }  // tuning override
        tc_grid_full = std::min(sms * bps, MAX_GRID);
        mma_flops = sup ? 3.0 * 2.0 * 128.0 * 192.0 : 3.0 * 2.0 * tc_padded_n(rows) * (64.0 * ((rows + 64) / 64));   // 3 passes over the padded NOUT x K tile (K covers N + 1 columns)
        tl.key0 = key0; tl.key1 = key1; tl.k = k; tl.d_image = e->d_tc_image;
        tl.flag_list = e->d_flag_list; tl.flag_count = e->d_flag_count; tl.flag_cap = FLAG_CAP;
        // (supplied normals: exact_pairs_kernel re-derives its normals from Philox, so every near-threshold trial takes the fp64 kernel)
        tl.pair_list = plan.recheck ? e->d_pair_list : nullptr; tl.pair_count = e->d_pair_count; tl.pair_cap = PAIR_CAP;
        tl.block_partials = e->d_block_partials;
    }

    // ---- row-slab tensor path (N > 128): per-slab images, per-trial uint16 accumulator ----
    const uint64_t chunk_paths = (plan.slab || plan.wide || plan.stream) ? SLAB_CHUNK_PATHS : CHUNK_PATHS;
    std::vector<double> slab_drel, slab_dabs;
    if (plan.slab) {
        const int ns = slab_count(N);
        const size_t bytes = slab_image_bytes(N, e->lower);
        if (e->slab_image_cap < bytes) {
            cudaFree(e->d_slab_image); e->d_slab_image = nullptr; e->slab_image_cap = 0;
            MCR_CUDA(cudaMalloc(&e->d_slab_image, bytes));
            e->slab_image_cap = bytes; e->slab_gen = 0;
        }
        if (!e->d_S_acc) MCR_CUDA(cudaMalloc(&e->d_S_acc, sizeof(uint16_t) * SLAB_CHUNK_PATHS));
        std::vector<double> c(N);
        for (int i = 0; i < N; ++i) c[i] = e->mu[i] + x;
        if (e->slab_gen != e->gen || e->slab_x != x) {
            e->slab_host_image.resize(bytes);
            e->slab_offsets.assign(ns + 1, 0);
            e->slab_row_scale.assign(N, 1.0);
            slab_build(N, e->lower, e->A.data(), c.data(), e->slab_host_image.data(), e->slab_offsets.data(), e->slab_row_scale.data());
            MCR_CUDA(cudaStreamSynchronize(s));
            MCR_CUDA(cudaMemcpyAsync(e->d_slab_image, e->slab_host_image.data(), bytes, cudaMemcpyHostToDevice, s));
            e->slab_gen = e->gen; e->slab_x = x;
        }
        // guard band per slab: fp16x3 split (3 * 2^-22) + fp32 accumulation over 12 MMAs per K-block
        static const double band_scale = [] { const char* v = getenv("MCR_TC_BAND_SCALE"); return v && *v ? atof(v) : 1.0; }();
        for (int t = 0; t < ns; ++t) {
            const int kbn = (slab_ncols(N, t, e->lower) + 1 + 63) / 64;
            // accumulation grows with the number of K-blocks; calibrated at N = 600 / 1024: counts stay exact with the
            // band at 1/16 of this setting (profiles/r01_tc_band_calibrate_slab_*.log)
            const double gamma = std::ldexp(1.0, -19) * std::max(1.0, kbn / 8.0);
            // sum_j |a_ij||z_j| <= ||a_i||_2 ||z||_2 (Cauchy-Schwarz): the kernel tracks ||z||_2 of the trial
            double cabs = 0, rownorm = 0;   // in accumulator units: row i is scaled by slab_row_scale[i]
            for (int i = slab_rows() * t; i < std::min(N, slab_rows() * (t + 1)); ++i) {
                cabs = std::max(cabs, std::fabs(c[i]) * e->slab_row_scale[i]);
                double ss = 0;
                for (int j = 0; j < N; ++j) ss += e->A[(size_t)i * N + j] * e->A[(size_t)i * N + j];
                rownorm = std::max(rownorm, std::sqrt(ss) * e->slab_row_scale[i]);
            }
            mma_flops += 3.0 * 2.0 * slab_rows() * 64.0 * kbn;
            slab_drel.push_back(plan.recheck ? band_scale * 2.0 * gamma * rownorm : 0.0);
            slab_dabs.push_back(plan.recheck ? band_scale * 2.0 * gamma * cabs + 1e-30 : 0.0);
        }
    }

    // ---- single-pass wide tensor path (128 < N <= 1024): block image, thresholds, per-slab guard bands ----
    WideLaunch wl{};
    if (plan.wide) {
        const int ns = wide_slabs(N), kbn = wide_kblocks(N);
        const size_t bytes = wide_image_bytes(N, e->lower);
        if (e->wide_image_cap < bytes) {
            cudaFree(e->d_wide_image); cudaFree(e->d_wide_c); e->d_wide_image = nullptr; e->d_wide_c = nullptr; e->wide_image_cap = 0;
            MCR_CUDA(cudaMalloc(&e->d_wide_image, bytes));
            MCR_CUDA(cudaMalloc(&e->d_wide_c, sizeof(float) * 128 * WIDE_MAX_SLABS));
            e->wide_image_cap = bytes; e->wide_gen = 0;
        }
        std::vector<double> c(N);
        for (int i = 0; i < N; ++i) c[i] = e->mu[i] + x;
        if (e->wide_gen != e->gen || e->wide_x != x) {
            e->wide_host_image.resize(bytes);
            e->wide_host_c.assign(128 * WIDE_MAX_SLABS, 0.f);
            e->wide_row_scale.assign(N, 1.0);
            wide_build(N, e->lower, e->A.data(), c.data(), e->wide_host_image.data(), e->wide_host_c.data(), e->wide_row_scale.data());
            MCR_CUDA(cudaStreamSynchronize(s));   // a previous run may still read the old image
            MCR_CUDA(cudaMemcpyAsync(e->d_wide_image, e->wide_host_image.data(), bytes, cudaMemcpyHostToDevice, s));
            MCR_CUDA(cudaMemcpyAsync(e->d_wide_c, e->wide_host_c.data(), sizeof(float) * 128 * ns, cudaMemcpyHostToDevice, s));
            e->wide_gen = e->gen; e->wide_x = x;
        }
        // guard band per 128-row slab, same budget as the row-slab kernel: fp16x3 split (3 * 2^-22) + fp32 accumulation over
        // 12 MMAs per K-block, relative to ||a_i||_2 ||z||_2 (Cauchy-Schwarz; the kernel uses the norm of all N normals)
        static const double band_scale = [] { const char* v = getenv("MCR_TC_BAND_SCALE"); return v && *v ? atof(v) : 1.0; }();
        for (int t = 0; t < ns; ++t) {
            const int kbt = e->lower ? std::min(2 * t + 2, kbn) : kbn;
            const double gamma = std::ldexp(1.0, -19) * std::max(1.0, kbt / 8.0);
            double rownorm = 0, cabs = 1e-300;
            for (int i = 128 * t; i < std::min(N, 128 * (t + 1)); ++i) {
                double ss = 0;
                for (int j = 0; j < N; ++j) ss += e->A[(size_t)i * N + j] * e->A[(size_t)i * N + j];
                rownorm = std::max(rownorm, std::sqrt(ss) * e->wide_row_scale[i]);
                cabs = std::max(cabs, std::fabs(c[i]) * e->wide_row_scale[i]);
            }
            wl.delta_rel[t] = plan.recheck ? band_scale * 2.0 * gamma * rownorm : 0.0;
            wl.delta_abs[t] = plan.recheck ? band_scale * 2.0 * gamma * cabs + 1e-30 : 0.0;
        }
        mma_flops = 3.0 * 2.0 * 128.0 * 64.0 * (double)(bytes / (2 * 128 * 128));   // per transform block: 128 rows x 64 columns, 3 passes
        wl.key0 = key0; wl.key1 = key1; wl.N = N; wl.k = k; wl.lower = e->lower ? 1 : 0;
        wl.d_image = e->d_wide_image; wl.d_c_scaled = e->d_wide_c;
        wl.flag_list = e->d_flag_list; wl.flag_count = e->d_flag_count; wl.flag_cap = FLAG_CAP;
        wl.pair_list = plan.recheck ? e->d_pair_list : nullptr; wl.pair_count = e->d_pair_count; wl.pair_cap = PAIR_CAP;
        wl.block_partials = e->d_block_partials;
    }

    // ---- streamed tensor path: piece image, per-pass guard bands ----
    StreamLaunch stl{};
    if (plan.stream) {
        const int port = plan.portfolio ? 1 : 0;
        const size_t bytes = stream_image_bytes(N, port, e->lower);
        if (e->stream_image_cap < bytes) {
            cudaFree(e->d_stream_image); e->d_stream_image = nullptr; e->stream_image_cap = 0;
            MCR_CUDA(cudaMalloc(&e->d_stream_image, bytes));
            e->stream_image_cap = bytes; e->stream_gen = 0;
        }
        std::vector<double> c(N), port_row;
        for (int i = 0; i < N; ++i) c[i] = e->mu[i] + x;
        if (port) {
            if (prepare_portfolio(port_row, stl.hist_lo, stl.hist_width)) return -1;
            stl.portfolio = 1;
            stl.port_x = opt->portfolio_loss_threshold;
            stl.port_acc = e->d_port_acc; stl.port_hist = e->d_port_hist; stl.port_hist_sum = e->d_port_hist_sum;
        }
        const bool pair = stream_pair(dZ != nullptr, port != 0);
        stl.pair = pair ? 1 : 0;
        const uint64_t image_key = e->gen * 8 + (uint64_t)port + (dZ ? 2 : 0) + (pair ? 4 : 0);   // (the piece order differs between generated and supplied normals)
        if (e->stream_gen != image_key || e->stream_x != x || port) {   // (the weights may have changed: a portfolio image is rebuilt)
            e->stream_host_image.resize(bytes);
            e->stream_row_scale.assign(N + port, 1.0);
            stream_build(N, e->lower, e->A.data(), c.data(), port ? port_row.data() : nullptr, dZ != nullptr, pair, e->stream_host_image.data(), e->stream_row_scale.data());
            MCR_CUDA(cudaStreamSynchronize(s));   // a previous run may still read the old image
            MCR_CUDA(cudaMemcpyAsync(e->d_stream_image, e->stream_host_image.data(), bytes, cudaMemcpyHostToDevice, s));
            e->stream_gen = image_key; e->stream_x = x;
        }
        if (port) stl.port_scale = e->stream_row_scale[N];
        // guard band per pass, same budget as the round-1 large-N kernels: fp16x3 split (3 * 2^-22) + fp32 accumulation over 12 MMAs
        // per K-block, relative to ||a_i||_2 ||z||_2 (Cauchy-Schwarz; the kernel tracks the norm of the pass's normals), in
        // accumulator units (row i scaled by stream_row_scale[i]); derivation in DESIGN.md section 3.1
        static const double band_scale = [] { const char* v = getenv("MCR_TC_BAND_SCALE"); return v && *v ? atof(v) : 1.0; }();
        StreamPass ps[STREAM_MAX_PASSES];
        const int np = stream_plan(N, port, e->lower, ps);
        for (int q = 0; q < np; ++q) {
            const double gamma = std::ldexp(1.0, -19) * std::max(1.0, ps[q].kbn / 8.0);
            double rownorm = 0, cabs = 1e-300;
            for (int i = ps[q].r0; i < std::min(N, ps[q].r0 + ps[q].W); ++i) {
                double ss = 0;
                for (int j = 0; j < N; ++j) ss += e->A[(size_t)i * N + j] * e->A[(size_t)i * N + j];
                rownorm = std::max(rownorm, std::sqrt(ss) * e->stream_row_scale[i]);
                cabs = std::max(cabs, std::fabs(c[i]) * e->stream_row_scale[i]);
            }
            stl.delta_rel[q] = plan.recheck ? band_scale * 2.0 * gamma * rownorm : 0.0;
            stl.delta_abs[q] = plan.recheck ? band_scale * 2.0 * gamma * cabs + 1e-30 : 0.0;
        }
        mma_flops = stream_mma_flops(N, port, e->lower);
        stl.key0 = key0; stl.key1 = key1; stl.N = N; stl.k = k; stl.lower = e->lower ? 1 : 0;
        stl.d_image = e->d_stream_image;
        stl.flag_list = e->d_flag_list; stl.flag_count = e->d_flag_count; stl.flag_cap = FLAG_CAP;
        // (supplied normals: exact_pairs_kernel reads the caller's matrix like the whole-trial kernel does)
        stl.pair_list = plan.recheck ? e->d_pair_list : nullptr; stl.pair_count = e->d_pair_count; stl.pair_cap = PAIR_CAP;
        stl.block_partials = e->d_block_partials;
    }

    for (uint64_t done = 0; done < M || (M == 0 && done == 0); done += chunk_paths) {
        const uint64_t m = std::min<uint64_t>(chunk_paths, M - done);
        if (m == 0) break;
        MCR_CUDA(cudaMemsetAsync(e->d_accum, 0, sizeof(unsigned long long) * 2, s));
        MCR_CUDA(cudaMemsetAsync(e->d_flag_count, 0, sizeof(uint32_t), s));
        MCR_CUDA(cudaMemsetAsync(e->d_pair_count, 0, sizeof(uint32_t), s));
        FinalizeParams fin{};
        fin.M = m; fin.out = d_out; fin.accum = e->d_accum;
        ExactParams xp{};
        xp.M = m; xp.path_offset = path_offset + done; xp.key0 = key0; xp.key1 = key1;
        xp.N = N; xp.k = k; xp.lower = e->lower ? 1 : 0; xp.x = x;
        xp.mu = e->d_mu; xp.A_cm = e->d_Acm;
        xp.Z = dZ ? dZ + done * (uint64_t)N : nullptr;
        xp.S_out = d_S ? (S_is_scratch ? d_S : d_S + done) : nullptr;
        xp.accum = e->d_accum;
        if (plan.slab) {
            const int ns = slab_count(N);
            for (int t = 0; t < ns; ++t) {
                SlabLaunch sl{};
                sl.M = m; sl.path_offset = path_offset + done; sl.key0 = key0; sl.key1 = key1;
                sl.ncols = slab_ncols(N, t, e->lower);
                sl.delta_rel = slab_drel[t]; sl.delta_abs = slab_dabs[t];
                sl.d_slab_image = (const uint8_t*)e->d_slab_image + e->slab_offsets[t];
                sl.S_acc = e->d_S_acc; sl.first_slab = t == 0;
                sl.row_base = slab_rows() * t;
                sl.pair_list = plan.recheck ? e->d_pair_list : nullptr; sl.pair_count = e->d_pair_count; sl.pair_cap = PAIR_CAP;
                const bool timed = e->n_timed < MAX_TIMED_CHUNKS;
                if (timed) MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed], s));
                MCR_CUDA(slab_launch(sl, e->num_sms, s));
                if (timed) { MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed + 1], s)); ++e->n_timed; }
                ++launches;
            }
            if (plan.recheck) {   // single near-threshold decisions: one fp64 row each, added to S_acc before it is counted
                PairParams pp{};
                pp.pair_list = e->d_pair_list; pp.pair_count = e->d_pair_count; pp.pair_cap = PAIR_CAP;
                pp.path_offset = path_offset + done; pp.key0 = key0; pp.key1 = key1;
                pp.N = N; pp.lower = e->lower ? 1 : 0; pp.x = x; pp.mu = e->d_mu; pp.A_rm = e->d_Arm; pp.S_acc = e->d_S_acc; pp.mode = 0;
                MCR_CUDA(launch_exact_pairs(pp, e->num_sms, s));
                ++launches;
            }
            const int grid = (int)std::min<uint64_t>((uint64_t)std::min(e->num_sms * 8, MAX_GRID), (m + 255) / 256);
            MCR_CUDA(slab_count_launch(m, e->d_S_acc, k, xp.S_out, e->d_flag_list, e->d_flag_count, FLAG_CAP, e->d_block_partials, grid, s));
            ++launches;
            if (plan.recheck) {
                xp.flag_list = e->d_flag_list; xp.flag_count = e->d_flag_count; xp.flag_cap = FLAG_CAP;
                MCR_CUDA(launch_exact(xp, e->num_sms, s));
                ++launches;
            }
            fin.block_partials = e->d_block_partials; fin.n_blocks = grid;
            fin.flag_count = e->d_flag_count; fin.flag_cap = FLAG_CAP;
            if (plan.recheck) { fin.pair_count = e->d_pair_count; fin.pair_cap = PAIR_CAP; }
        } else if (plan.stream) {
            stl.M = m; stl.path_offset = path_offset + done; stl.S_out = xp.S_out; stl.d_Z = xp.Z;
            const bool timed = e->n_timed < MAX_TIMED_CHUNKS;
            if (timed) MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed], s));
            MCR_CUDA(stream_launch(stl, sms, s));
            if (timed) { MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed + 1], s)); ++e->n_timed; }
            ++launches;
            if (plan.recheck) {
                if (stl.pair_list) {
                    PairParams pp{};
                    pp.pair_list = e->d_pair_list; pp.pair_count = e->d_pair_count; pp.pair_cap = PAIR_CAP;
                    pp.path_offset = path_offset + done; pp.key0 = key0; pp.key1 = key1;
                    pp.N = N; pp.lower = e->lower ? 1 : 0; pp.x = x; pp.mu = e->d_mu; pp.A_rm = e->d_Arm;
                    pp.k = k; pp.S_out = xp.S_out; pp.accum = e->d_accum; pp.mode = 2; pp.Z = xp.Z;
                    MCR_CUDA(launch_exact_pairs(pp, e->num_sms, s));
                    ++launches;
                    fin.pair_count = e->d_pair_count; fin.pair_cap = PAIR_CAP;
                }
                xp.flag_list = e->d_flag_list; xp.flag_count = e->d_flag_count; xp.flag_cap = FLAG_CAP;
                MCR_CUDA(launch_exact(xp, e->num_sms, s));
                ++launches;
            }
            fin.block_partials = e->d_block_partials; fin.n_blocks = stream_grid(m, sms, stl.pair != 0);
            fin.flag_count = e->d_flag_count; fin.flag_cap = FLAG_CAP;
        } else if (plan.wide) {
            wl.M = m; wl.path_offset = path_offset + done; wl.S_out = xp.S_out;
            const bool timed = e->n_timed < MAX_TIMED_CHUNKS;
            if (timed) MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed], s));
            MCR_CUDA(wide_launch(wl, e->num_sms, s));
            if (timed) { MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed + 1], s)); ++e->n_timed; }
            ++launches;
            if (plan.recheck) {
                PairParams pp{};
                pp.pair_list = e->d_pair_list; pp.pair_count = e->d_pair_count; pp.pair_cap = PAIR_CAP;
                pp.path_offset = path_offset + done; pp.key0 = key0; pp.key1 = key1;
                pp.N = N; pp.lower = e->lower ? 1 : 0; pp.x = x; pp.mu = e->d_mu; pp.A_rm = e->d_Arm;
                pp.k = k; pp.S_out = xp.S_out; pp.accum = e->d_accum; pp.mode = 2;
                MCR_CUDA(launch_exact_pairs(pp, e->num_sms, s));
                ++launches;
                fin.pair_count = e->d_pair_count; fin.pair_cap = PAIR_CAP;
                xp.flag_list = e->d_flag_list; xp.flag_count = e->d_flag_count; xp.flag_cap = FLAG_CAP;
                MCR_CUDA(launch_exact(xp, e->num_sms, s));
                ++launches;
            }
            fin.block_partials = e->d_block_partials; fin.n_blocks = wide_grid(m, N, e->num_sms);
            fin.flag_count = e->d_flag_count; fin.flag_cap = FLAG_CAP;
        } else if (plan.kernel == MCR_KERNEL_TENSOR) {
            tl.M = m; tl.path_offset = path_offset + done; tl.S_out = xp.S_out; tl.d_Z = xp.Z;
            const uint64_t tiles = (m + 127) / 128;
            const int grid = (int)std::min<uint64_t>((uint64_t)tc_grid_full, tiles);
            const bool timed = e->n_timed < MAX_TIMED_CHUNKS;
            if (timed) MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed], s));
            MCR_CUDA(tc_launch(N, plan.rows, tl, grid, s));
            if (timed) { MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed + 1], s)); ++e->n_timed; }
            ++launches;
            if (plan.recheck) {
                PairParams pp{};   // trials with a single near-threshold decision: that row alone in fp64
                pp.pair_list = e->d_pair_list; pp.pair_count = e->d_pair_count; pp.pair_cap = PAIR_CAP;
                pp.path_offset = path_offset + done; pp.key0 = key0; pp.key1 = key1;
                pp.N = N; pp.lower = e->lower ? 1 : 0; pp.x = x; pp.mu = e->d_mu; pp.A_rm = e->d_Arm;
                pp.k = k; pp.S_out = xp.S_out; pp.accum = e->d_accum; pp.mode = 1; pp.Z = xp.Z;
                MCR_CUDA(launch_exact_pairs(pp, e->num_sms, s));
                ++launches;
                fin.pair_count = e->d_pair_count; fin.pair_cap = PAIR_CAP;
                xp.flag_list = e->d_flag_list; xp.flag_count = e->d_flag_count; xp.flag_cap = FLAG_CAP;
                MCR_CUDA(launch_exact(xp, e->num_sms, s));
                ++launches;
            }
            fin.block_partials = e->d_block_partials; fin.n_blocks = grid;
            fin.flag_count = e->d_flag_count; fin.flag_cap = FLAG_CAP;
        } else if (var) {
            fp.M = m; fp.path_offset = path_offset + done;
            fp.Z = xp.Z; fp.S_out = xp.S_out;
            const uint64_t need = (m + 127) / 128;
            const int grid = (int)std::min<uint64_t>((uint64_t)grid_full, need);
            const bool timed = e->n_timed < MAX_TIMED_CHUNKS;
            if (timed) MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed], s));
            MCR_CUDA(var->launch(fp, grid, 128, s));
            if (timed) { MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed + 1], s)); ++e->n_timed; }
            ++launches;
            if (plan.recheck) {
                xp.flag_list = e->d_flag_list; xp.flag_count = e->d_flag_count; xp.flag_cap = FLAG_CAP;
                MCR_CUDA(launch_exact(xp, e->num_sms, s));
                ++launches;
            }
            fin.block_partials = e->d_block_partials; fin.n_blocks = grid;
            fin.flag_count = e->d_flag_count; fin.flag_cap = FLAG_CAP;
        } else {
            const bool timed = e->n_timed < MAX_TIMED_CHUNKS;
            if (timed) MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed], s));
            MCR_CUDA(launch_exact(xp, e->num_sms, s));
            if (timed) { MCR_CUDA(cudaEventRecord(e->evk[2 * e->n_timed + 1], s)); ++e->n_timed; }
            ++launches;
            fin.block_partials = nullptr; fin.n_blocks = 0; fin.flag_count = nullptr;
        }
        MCR_CUDA(launch_finalize(fin, s));
        ++launches;
        if (want_hist) {
            MCR_CUDA(launch_hist(m, xp.S_out, N + 1, e->d_hist, e->num_sms, s));
            ++launches;
        }
    }
    e->last_kernel = plan.kernel; e->last_NT = plan.NT; e->last_launches = (accumulate ? e->last_launches : 0) + launches; e->last_portfolio = plan.portfolio;
    e->last_mma_flops = mma_flops;
    MCR_CUDA(cudaEventRecord(e->ev_run_done, s));
    e->run_pending = true;
    return 0;
}

// extension accumulators -> the two contiguous partial buffers of mcr_run_async_ext (zeros where not requested)
__global__ void gather_ext_kernel(const unsigned long long* hist, const double* port_acc, const unsigned long long* port_hist,
                                  const double* port_hist_sum, int have_hist, int have_port, unsigned long long* counts, double* sums) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < MCR_MAX_HIST) counts[t] = have_hist ? hist[t] : 0ull;
    if (t == 0) counts[MCR_MAX_HIST] = have_port ? (unsigned long long)(port_acc[0] + 0.5) : 0ull;
    if (t < 3) sums[t] = have_port ? port_acc[1 + t] : 0.0;
    if (t < MCR_PORTFOLIO_BINS) {
        counts[MCR_MAX_HIST + 1 + t] = have_port ? port_hist[t] : 0ull;
        sums[3 + t] = have_port ? port_hist_sum[t] : 0.0;
    }
}

int enqueue_gather_ext(mcr_engine* e, uint64_t* d_ext_counts, double* d_ext_sums, cudaStream_t s) {
    static_assert(MCR_PORTFOLIO_BINS >= MCR_MAX_HIST, "gather_ext_kernel: one thread per portfolio bin covers hist_S too");
    gather_ext_kernel<<<(MCR_PORTFOLIO_BINS + 255) / 256, 256, 0, s>>>(
        e->d_hist, e->d_port_acc, e->d_port_hist, e->d_port_hist_sum, e->last_hist ? 1 : 0, e->last_portfolio ? 1 : 0,
        reinterpret_cast<unsigned long long*>(d_ext_counts), d_ext_sums);
    MCR_CUDA(cudaGetLastError());
    ++e->last_launches;
    return 0;
}

void fill_stats(mcr_result* r) {
    const double M = (double)r->n_paths;
    if (r->n_paths == 0) { r->P_hat = r->std_error = r->ci_lower = r->ci_upper = 0.0; return; }
    r->P_hat = (double)r->count / M;
    r->std_error = std::sqrt(r->P_hat * (1.0 - r->P_hat) / M);
    const double margin = 1.96 * r->std_error;
    r->ci_lower = std::fmax(0.0, r->P_hat - margin);
    r->ci_upper = std::fmin(1.0, r->P_hat + margin);
}

}  // namespace mcr

extern "C" {

int mcr_last_main_kernel_ms(mcr_engine* e, float* avg_ms, int* n_launches) {
    if (!e || !avg_ms || !n_launches) return fail("mcr_last_main_kernel_ms: null argument");
    if (set_device(e)) return -1;
    double sum = 0;
    for (int i = 0; i < e->n_timed; ++i) {
        MCR_CUDA(cudaEventSynchronize(e->evk[2 * i + 1]));
        float ms = 0;
        MCR_CUDA(cudaEventElapsedTime(&ms, e->evk[2 * i], e->evk[2 * i + 1]));
        sum += ms;
    }
    *n_launches = e->n_timed;
    *avg_ms = e->n_timed ? (float)(sum / e->n_timed) : 0.f;
    return 0;
}

int mcr_last_mma_flops_per_trial(mcr_engine* e, double* flops) {
    if (!e || !flops) return fail("mcr_last_mma_flops_per_trial: null argument");
    *flops = e->last_mma_flops;
    return 0;
}

int mcr_result_from_partials(const uint64_t* h, mcr_result* out) {
    if (!h || !out) return fail("mcr_result_from_partials: null argument");
    memset(out, 0, sizeof(*out));
    out->n_paths = h[0]; out->count = h[1]; out->sum_S = h[2]; out->n_rechecked = h[3];
    fill_stats(out);
    if (h[4] != 0) return fail("fix-up list overflowed (more than %u near-threshold trials in one chunk); result incomplete", FLAG_CAP);
    return 0;
}

int mcr_run_async(mcr_engine* e, uint64_t M, uint64_t path_offset, uint64_t seed, double x, int k,
                  const mcr_run_options* opt, uint64_t* d_partials, void* cuda_stream) {
    if (!e || !d_partials) return fail("mcr_run_async: null argument");
    if (opt && (opt->want_hist || opt->want_portfolio)) return fail("mcr_run_async: extension outputs need mcr_run_async_ext (or mcr_run)");
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : e->stream;
    return enqueue_run(e, M, path_offset, seed, x, k, opt, nullptr, d_partials, s);
}

int mcr_run_async_ext(mcr_engine* e, uint64_t M, uint64_t path_offset, uint64_t seed, double x, int k,
                      const mcr_run_options* opt, uint64_t* d_partials, uint64_t* d_ext_counts, double* d_ext_sums, void* cuda_stream) {
    if (!e || !d_partials || !d_ext_counts || !d_ext_sums) return fail("mcr_run_async_ext: null argument");
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : e->stream;
    if (enqueue_run(e, M, path_offset, seed, x, k, opt, nullptr, d_partials, s)) return -1;
    return enqueue_gather_ext(e, d_ext_counts, d_ext_sums, s);
}

int mcr_portfolio_hist_range(mcr_engine* e, double* lo, double* hi) {
    if (!e || !lo || !hi) return fail("mcr_portfolio_hist_range: null argument");
    if (!e->last_portfolio) return fail("mcr_portfolio_hist_range: the engine's last run had no want_portfolio");
    *lo = e->port_hist_lo; *hi = e->port_hist_hi;
    return 0;
}

int mcr_result_from_partials_ext(const uint64_t* h, const uint64_t* hc, const double* hs, double hist_lo, double hist_hi, mcr_result* out) {
    if (!hc || !hs) return fail("mcr_result_from_partials_ext: null argument");
    const int rc = mcr_result_from_partials(h, out);
    memcpy(out->hist_S, hc, sizeof(uint64_t) * MCR_MAX_HIST);
    out->portfolio_hits = hc[MCR_MAX_HIST];
    memcpy(out->portfolio_hist, hc + MCR_MAX_HIST + 1, sizeof(uint64_t) * MCR_PORTFOLIO_BINS);
    out->portfolio_sum = hs[0]; out->portfolio_sumsq = hs[1]; out->portfolio_tail_loss_sum = hs[2];
    memcpy(out->portfolio_hist_sum, hs + 3, sizeof(double) * MCR_PORTFOLIO_BINS);
    out->portfolio_hist_lo = hist_lo; out->portfolio_hist_hi = hist_hi;
    return rc;
}

static int finish_sync(mcr_engine* e, mcr_result* out) {
    MCR_CUDA(cudaEventRecord(e->ev1, e->stream));
    MCR_CUDA(cudaMemcpyAsync(e->h_partials, e->d_partials, sizeof(uint64_t) * MCR_PARTIAL_WORDS, cudaMemcpyDeviceToHost, e->stream));
    MCR_CUDA(cudaStreamSynchronize(e->stream));
    const int rc = mcr_result_from_partials(e->h_partials, out);
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    out->device_ms = ms;
    out->kernel_used = e->last_kernel; out->padded_N = e->last_NT; out->launches = e->last_launches;
    return rc;
}

int mcr_run(mcr_engine* e, uint64_t M, uint64_t path_offset, uint64_t seed, double x, int k,
            const mcr_run_options* opt, mcr_result* out) {
    if (!e || !out) return fail("mcr_run: null argument");
    if (set_device(e)) return -1;
    MCR_CUDA(cudaEventRecord(e->ev0, e->stream));
    if (enqueue_run(e, M, path_offset, seed, x, k, opt, nullptr, e->d_partials, e->stream)) return -1;
    const int rc = finish_sync(e, out);
    if (rc == 0 && e->last_hist)
        MCR_CUDA(cudaMemcpy(out->hist_S, e->d_hist, sizeof(uint64_t) * MCR_MAX_HIST, cudaMemcpyDeviceToHost));
    if (rc == 0 && e->last_portfolio) {
        double acc[4];
        MCR_CUDA(cudaMemcpy(acc, e->d_port_acc, sizeof(acc), cudaMemcpyDeviceToHost));
        MCR_CUDA(cudaMemcpy(out->portfolio_hist, e->d_port_hist, sizeof(uint64_t) * MCR_PORTFOLIO_BINS, cudaMemcpyDeviceToHost));
        MCR_CUDA(cudaMemcpy(out->portfolio_hist_sum, e->d_port_hist_sum, sizeof(double) * MCR_PORTFOLIO_BINS, cudaMemcpyDeviceToHost));
        out->portfolio_hits = (uint64_t)(acc[0] + 0.5);
        out->portfolio_sum = acc[1]; out->portfolio_sumsq = acc[2]; out->portfolio_tail_loss_sum = acc[3];
        out->portfolio_hist_lo = e->port_hist_lo; out->portfolio_hist_hi = e->port_hist_hi;
    }
    return rc;
}

int mcr_run_with_normals_device(mcr_engine* e, uint64_t M, const double* d_Z, double x, int k,
                                const mcr_run_options* opt, mcr_result* out) {
    if (!e || !out || (!d_Z && M)) return fail("mcr_run_with_normals_device: null argument");
    if (set_device(e)) return -1;
    MCR_CUDA(cudaEventRecord(e->ev0, e->stream));
    if (enqueue_run(e, M, 0, 0, x, k, opt, d_Z, e->d_partials, e->stream)) return -1;
    return finish_sync(e, out);
}

int mcr_run_with_normals(mcr_engine* e, uint64_t M, const double* Z, double x, int k,
                         const mcr_run_options* opt, int32_t* S_values_host, mcr_result* out) {
    if (!e || !out || (!Z && M)) return fail("mcr_run_with_normals: null argument");
    if (e->N <= 0) return fail("no model uploaded (call mcr_upload_model first)");
    if (set_device(e)) return -1;
    double* dZ = nullptr;
    int32_t* dS = nullptr;
    const size_t zbytes = sizeof(double) * (size_t)M * e->N;
    if (M) MCR_CUDA(cudaMalloc(&dZ, zbytes));
    int rc = -1;
    do {
        if (M && cudaMemcpy(dZ, Z, zbytes, cudaMemcpyHostToDevice) != cudaSuccess) { fail("H2D copy of normals failed"); break; }
        mcr_run_options o{};
        if (opt) o = *opt;
        o.struct_size = sizeof(o);
        if (S_values_host && M) {
            if (cudaMalloc(&dS, sizeof(int32_t) * M) != cudaSuccess) { fail("S_values allocation failed"); break; }
            o.d_S_values = dS;
        }
        rc = mcr_run_with_normals_device(e, M, dZ, x, k, &o, out);
        if (rc == 0 && dS && cudaMemcpy(S_values_host, dS, sizeof(int32_t) * M, cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = fail("D2H copy of S_values failed");
    } while (0);
    cudaFree(dZ); cudaFree(dS);
    return rc;
}

int mcr_dump_normals(mcr_engine* e, uint64_t M, uint64_t path_offset, uint64_t seed, float* Z_host, uint32_t* U_host) {
    if (!e) return fail("mcr_dump_normals: null engine");
    if (e->N <= 0) return fail("no model uploaded (call mcr_upload_model first)");
    if (set_device(e)) return -1;
    const int N = e->N, nb = (N + 3) / 4;
    float* dZ = nullptr; uint32_t* dU = nullptr;
    if (Z_host) MCR_CUDA(cudaMalloc(&dZ, sizeof(float) * M * N));
    if (U_host && cudaMalloc(&dU, sizeof(uint32_t) * M * nb * 4) != cudaSuccess) { cudaFree(dZ); return fail("allocation failed"); }
    int rc = 0;
    if (launch_dump(M, path_offset, (uint32_t)seed, (uint32_t)(seed >> 32), N, dZ, dU, e->stream) != cudaSuccess ||
        cudaStreamSynchronize(e->stream) != cudaSuccess)
        rc = fail("dump kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (!rc && Z_host && cudaMemcpy(Z_host, dZ, sizeof(float) * M * N, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail("D2H failed");
    if (!rc && U_host && cudaMemcpy(U_host, dU, sizeof(uint32_t) * M * nb * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail("D2H failed");
    cudaFree(dZ); cudaFree(dU);
    return rc;
}

int mcr_portfolio_var_es(const mcr_result* r, double alpha, double* var_out, double* es_out) {
    if (!r || !var_out || !es_out) return fail("mcr_portfolio_var_es: null argument");
    if (!(alpha > 0.0 && alpha < 1.0)) return fail("mcr_portfolio_var_es: alpha must be in (0, 1)");
    uint64_t total = 0;
    for (int b = 0; b < MCR_PORTFOLIO_BINS; ++b) total += r->portfolio_hist[b];
    if (total == 0) return fail("mcr_portfolio_var_es: the result carries no portfolio histogram (run with want_portfolio)");
    // walk up from the lowest returns until (1 - alpha) of the trials are covered
    const double target = (1.0 - alpha) * (double)total;
    const double width = (r->portfolio_hist_hi - r->portfolio_hist_lo) / MCR_PORTFOLIO_BINS;
    double cum = 0, tail_sum = 0;
    for (int b = 0; b < MCR_PORTFOLIO_BINS; ++b) {
        const double c = (double)r->portfolio_hist[b];
        if (cum + c >= target && c > 0) {
            const double f = (target - cum) / c;   // fraction of this bin inside the tail
            const double p_var = r->portfolio_hist_lo + width * (b + f);
            // within the bin take the proportional share of its return sum
            tail_sum += f * r->portfolio_hist_sum[b];
            *var_out = -p_var;
            *es_out = target > 0 ? -(tail_sum / target) : -p_var;
            return 0;
        }
        cum += c;
        tail_sum += r->portfolio_hist_sum[b];
    }
    return fail("mcr_portfolio_var_es: histogram exhausted");
}

int mcr_measure_fp32_peak(mcr_engine* e, float ms_target, double* tflops_out) {
    if (!e || !tflops_out) return fail("mcr_measure_fp32_peak: null argument");
    if (set_device(e)) return -1;
    float* sink = nullptr;
    MCR_CUDA(cudaMalloc(&sink, sizeof(float)));
    const int grid = e->num_sms * 8;
    int iters = 2000;
    float ms = 0;
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e->ev0, e->stream);
        if (launch_fp32_peak(grid, iters, sink, e->stream) != cudaSuccess) { cudaFree(sink); return fail("fp32 peak kernel launch failed"); }
        cudaEventRecord(e->ev1, e->stream);
        if (cudaStreamSynchronize(e->stream) != cudaSuccess) { cudaFree(sink); return fail("fp32 peak kernel failed"); }
        cudaEventElapsedTime(&ms, e->ev0, e->ev1);
        const double flops = (double)grid * 256.0 * (double)iters * 8.0 * 16.0 * 4.0;  // 2 lanes x 2 flops per FFMA2
        if (rep >= 1) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        if (ms < ms_target && rep < 2) iters = (int)std::min(2.0e6, iters * (double)ms_target / std::max(ms, 0.01f));
    }
    cudaFree(sink);
    *tflops_out = best;
    return 0;
}

}  // extern "C"
