// Engine object behind the opaque mcr_engine handle (internal).
#pragma once
#include "../../include/mcr_b200.h"
#include "mcr_internal.h"

#include <vector>

namespace mcr {
constexpr uint64_t CHUNK_PATHS = 1ull << 27;   // trials per launch of the path kernel
constexpr uint64_t SLAB_CHUNK_PATHS = 1ull << 24;  // trials per chunk on the row-slab path (uint16 accumulator per trial)
constexpr uint32_t FLAG_CAP = 1u << 22;        // fix-up list entries per chunk (3.1 % of a chunk)
constexpr uint32_t PAIR_CAP = 1u << 22;        // slab path: single (trial, row) fix-ups per chunk of 2^24 trials
constexpr int MAX_GRID = 148 * 16 * 2;
constexpr int MAX_TIMED_CHUNKS = 16;
int fail(const char* fmt, ...);
}  // namespace mcr

#define MCR_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return mcr::fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct mcr_engine {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_run_done = nullptr;   // recorded at the end of every enqueued run: the next run (any stream) waits for it
    bool run_pending = false;
    bool accumulate = false;   // set by mcr_simulate_host between the parts of one run: the next enqueue_run adds to the outputs instead of zeroing them
    cudaEvent_t evk[2 * mcr::MAX_TIMED_CHUNKS] = {};  // (start, stop) around each path-kernel launch of the last run
    int n_timed = 0;
    // model
    int N = 0;
    int N_alloc = 0;
    bool lower = false;
    uint64_t gen = 0;
    std::vector<double> mu, A, w;
    std::vector<std::vector<float>> packed;  // per variant index, filled lazily
    double* d_mu = nullptr;
    double* d_Acm = nullptr;
    // tensor-core path: device copy of the split transform image and what it was built for
    void* d_tc_image = nullptr;
    size_t tc_image_cap = 0;
    uint64_t tc_gen = 0;
    double tc_x = 0;
    std::vector<uint8_t> tc_host_image;
    std::vector<double> tc_row_scale;     // power-of-two scale of every accumulator row of the current image
    // row-slab tensor path (N > 128)
    void* d_slab_image = nullptr;
    size_t slab_image_cap = 0;
    uint64_t slab_gen = 0;
    double slab_x = 0;
    std::vector<uint8_t> slab_host_image;
    std::vector<size_t> slab_offsets;
    std::vector<double> slab_row_scale;
    uint16_t* d_S_acc = nullptr;
    // single-pass wide tensor path (128 < N <= 1024)
    void* d_wide_image = nullptr;
    float* d_wide_c = nullptr;
    size_t wide_image_cap = 0;
    uint64_t wide_gen = 0;
    double wide_x = 0;
    std::vector<uint8_t> wide_host_image;
    std::vector<float> wide_host_c;
    std::vector<double> wide_row_scale;
    // streamed tensor path (N > 128 by default; any N with supplied normals)
    void* d_stream_image = nullptr;
    size_t stream_image_cap = 0;
    uint64_t stream_gen = 0;
    double stream_x = 0;
    std::vector<uint8_t> stream_host_image;
    std::vector<double> stream_row_scale;
    unsigned long long* d_pair_list = nullptr;   // single near-threshold (trial, row) decisions of a chunk
    uint32_t* d_pair_count = nullptr;
    double* d_Arm = nullptr;                     // fp64 transform, row-major (per-decision fix-up)
    double gamma_rowabs_max = 0;  // max_i gamma_i * sum_j |A_ij|
    // scratch
    uint32_t* d_flag_list = nullptr;
    uint32_t* d_flag_count = nullptr;
    unsigned long long* d_accum = nullptr;
    uint64_t* d_block_partials = nullptr;
    uint64_t* d_partials = nullptr;
    uint64_t* h_partials = nullptr;  // pinned
    // portfolio extension accumulators (device)
    double* d_port_acc = nullptr;
    unsigned long long* d_port_hist = nullptr;
    double* d_port_hist_sum = nullptr;
    double port_hist_lo = 0, port_hist_hi = 0;
    bool last_portfolio = false;
    // histogram of S (want_hist): device bins + a per-chunk S buffer when the caller supplied none
    unsigned long long* d_hist = nullptr;
    int32_t* d_S_tmp = nullptr;
    uint64_t S_tmp_cap = 0;
    bool last_hist = false;
    // S_values streaming (mcr_simulate_host, want_S_values): per-part int32 scratch, two compact device buffers, two pinned host
    // buffers, a copy stream and the events that order them
    int32_t* d_S_part = nullptr;
    uint8_t* d_S_compact[2] = {nullptr, nullptr};
    uint8_t* h_S_compact[2] = {nullptr, nullptr};
    uint64_t S_part_cap = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_compact[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    // extension partials in the mcr_run_async_ext layout (mcr_simulate_host; allocated on first use)
    uint64_t* d_ext_counts = nullptr;
    double* d_ext_sums = nullptr;
    std::vector<uint64_t> h_ext_counts;
    std::vector<double> h_ext_sums;
    // last async run bookkeeping
    int last_kernel = 0, last_NT = 0, last_launches = 0;
    double last_mma_flops = 0;   // fp16 MMA flops per trial the tensor pipe executed in the last run
};

namespace mcr {
// enqueue one full run (all chunks) on stream s; partial words accumulate into d_out (zeroed first).
// dZ != nullptr selects normals mode (device, M x N fp64).
int enqueue_run(mcr_engine* e, uint64_t M, uint64_t path_offset, uint64_t seed, double x, int k,
                const mcr_run_options* opt, const double* dZ, uint64_t* d_out, cudaStream_t s);
void fill_stats(mcr_result* r);
// copy the extension accumulators of the run just enqueued on s into the caller's contiguous buffers (device)
int enqueue_gather_ext(mcr_engine* e, uint64_t* d_ext_counts, double* d_ext_sums, cudaStream_t s);
}  // namespace mcr
