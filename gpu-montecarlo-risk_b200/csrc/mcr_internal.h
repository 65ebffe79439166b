// Internal declarations shared by the CUDA translation units of libmcr_b200.so.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mcr {

// ---- parameters of the fused FFMA path kernel (mcr_ffma.cuh) ------------------------------
struct FfmaParams {
    uint64_t M;             // paths of this launch (chunk)
    uint64_t path_offset;   // global index of the chunk's first path (Philox counter)
    uint32_t key0, key1;    // Philox key = seed
    int32_t k;              // tail-event threshold on S
    int32_t N;              // real asset count (<= NT); row length of Z in normals mode
    float delta;            // Philox mode: guard band on |R_i + x| (0 disables the recheck)
    float delta_rel;        // normals mode: band = delta_rel * max_j|z_j| + delta_abs
    float delta_abs;
    const double* Z;        // normals mode: device, M x N row-major fp64 (chunk-relative)
    int32_t* S_out;         // optional per-path S (chunk-relative), or nullptr
    uint32_t* flag_list;    // chunk-relative indices of paths that need the fp64 fix-up
    uint32_t* flag_count;
    uint32_t flag_cap;
    uint64_t* block_partials;  // [gridDim.x][2] = {count, sum_S} of the unflagged paths
};

// one compiled instantiation of the FFMA kernel
struct FfmaVariant {
    int NT;          // padded asset count the code is unrolled for
    bool lower;      // triangular (skips the strict upper triangle)
    bool normals;    // normals read from memory instead of generated
    int ppt;         // trials carried per thread (transform loads amortised over them)
    int n_pairs;     // float2 entries of the packed transform
    // pack A (N x N fp64 row-major) into this variant's constant layout (host)
    void (*pack)(const double* A, int N, float* packed /* 2*n_pairs floats */);
    // copy packed transform / thresholds into the variant's __constant__ storage
    cudaError_t (*upload_A)(const float* packed, cudaStream_t s);
    cudaError_t (*upload_c)(const float* c /* NT floats */, cudaStream_t s);
    cudaError_t (*launch)(const FfmaParams& p, int grid, int block, cudaStream_t s);
    cudaError_t (*occupancy)(int block, int* blocks_per_sm);
    const void* symbol_tag;  // identity of the __constant__ storage this variant shares
};

constexpr int MAX_VARIANTS = 128;
struct VariantRegistry {
    FfmaVariant v[MAX_VARIANTS];
    int n = 0;
};
VariantRegistry& registry();
void register_variant(const FfmaVariant& v);

// instantiation groups (one translation unit each, see mcr_ffma_inst.cu)
void register_group0(); void register_group1(); void register_group2(); void register_group3();
void register_group4(); void register_group5(); void register_group6(); void register_group7();

// ---- fp64 exact / fix-up kernel (mcr_exact.cu) ---------------------------------------------
struct ExactParams {
    uint64_t M;                 // paths of the chunk (used when flag_list == nullptr: all paths)
    uint64_t path_offset;
    uint32_t key0, key1;
    int32_t N, k;
    int32_t lower;              // skip the strict upper triangle (exact: those terms are +-0)
    double x;
    const double* mu;           // N
    const double* A_cm;         // N x N column-major fp64: A_cm[j*N + i] = A[i][j]
    const double* Z;            // normals mode (chunk-relative) or nullptr -> regenerate with Philox
    const uint32_t* flag_list;  // or nullptr
    const uint32_t* flag_count;
    uint32_t flag_cap;
    int32_t* S_out;
    unsigned long long* accum;  // [0] += #{S>=k}, [1] += sum S   (integer atomics: order-independent)
    unsigned long long* hist;   // optional S histogram (N+1 bins) or nullptr
};
cudaError_t launch_exact(const ExactParams& p, int num_sms, cudaStream_t s);

// fp64 re-evaluation of single (trial, row) decisions the slab kernel left out of S_acc (mcr_exact.cu)
struct PairParams {
    const unsigned long long* pair_list;   // mode 0 (slab path): trial << 16 | row;  1 (N <= 128): trial << 16 | S_rest << 8 | row;
                                           // 2 (single-pass wide kernel): trial << 21 | S_rest << 10 | row
    int32_t mode;
    const uint32_t* pair_count;
    uint32_t pair_cap;
    uint64_t path_offset;
    uint32_t key0, key1;
    int32_t N, lower;
    double x;
    const double* mu;
    const double* A_rm;                    // N x N row-major fp64
    const double* Z;                       // supplied normals (M x N fp64, chunk-relative like the trial indices) or nullptr: Philox
    uint16_t* S_acc;                       // mode 0: hit added here (two trials per 32-bit word)
    int32_t k;                             // modes 1, 2: S = S_rest + hit is final: counted into accum like exact_kernel does
    int32_t* S_out;
    unsigned long long* accum;
};
cudaError_t launch_exact_pairs(const PairParams& p, int num_sms, cudaStream_t s);

struct FinalizeParams {
    const uint64_t* block_partials;  // [n_blocks][2]
    int n_blocks;
    const unsigned long long* accum; // fix-up accumulators [2]
    const uint32_t* flag_count;
    uint32_t flag_cap;
    const uint32_t* pair_count;      // slab path: per-decision fix-ups (counted as rechecked), or nullptr
    uint32_t pair_cap;
    uint64_t M;
    uint64_t* out;                   // MCR_PARTIAL_WORDS, accumulated (+=)
};
cudaError_t launch_finalize(const FinalizeParams& p, cudaStream_t s);

// ---- tensor-core path (mcr_tc.cu) -----------------------------------------------------------------
struct TcLaunch {
    uint64_t M, path_offset;
    uint32_t key0, key1;
    int32_t k;
    double delta_rel, delta_abs; // guard band = delta_rel * max_j|z_j| + delta_abs, in accumulator units (row scales applied)
    const void* d_image;        // device copy of the split, swizzled transform image (tc_build / tc_supplied_build)
    const double* d_Z;          // supplied normals (M x N fp64, chunk-relative): the supplied-normals instantiation runs; else nullptr
    int32_t* S_out;
    uint32_t* flag_list;
    uint32_t* flag_count;
    uint32_t flag_cap;
    unsigned long long* pair_list;  // trial << 16 | S without that decision << 8 | row: single near-threshold decisions (or nullptr)
    uint32_t* pair_count;
    uint32_t pair_cap;
    uint64_t* block_partials;
    // portfolio extension
    int32_t portfolio;
    double port_x, hist_lo, hist_width;
    double port_scale;          // power-of-two scale of the portfolio row (tc_build's row_scale[rows - 1])
    double* port_acc;
    unsigned long long* port_hist;
    double* port_hist_sum;
};
int tc_max_n();
// `rows` = accumulator rows needed: N, or N + 1 when the last row carries the portfolio return
bool tc_supported(int rows);
size_t tc_image_bytes(int rows);
int tc_padded_n(int rows);
// A: N x N row-major fp64 transform, c: N thresholds mu_i + x, port_row: nullptr or N + 1 values
// (w^T A, then w.mu)  ->  host image (tc_image_bytes(rows))
// row_scale: out, tc_padded_n(rows) entries = the power-of-two scale of every accumulator row (tc_row_scale)
void tc_build(int N, int rows, const double* A, const double* c, const double* port_row, uint8_t* image, double* row_scale);
double tc_z_factor();   // what tc_build multiplies the transform columns by (the generator's folded sqrt(2 ln 2), or 1)
cudaError_t tc_occupancy(int rows, bool portfolio, int* blocks_per_sm);
// the supplied-normals instantiation of the tile kernel: one shape (128 rows, 3 K-blocks) for every N <= 128
size_t tc_supplied_image_bytes();
int tc_supplied_padded_n();
void tc_supplied_build(int N, const double* A, const double* c, uint8_t* image, double* row_scale /* 128, out */);
cudaError_t tc_supplied_occupancy(int* blocks_per_sm);
cudaError_t tc_launch(int N, int rows, const TcLaunch& l, int grid, cudaStream_t s);

// ---- tensor-core row-slab path for N > 128 (mcr_tc.cu) --------------------------------------------
struct SlabLaunch {
    uint64_t M, path_offset;
    uint32_t key0, key1;
    int32_t ncols;               // normal columns this slab multiplies
    double delta_rel, delta_abs; // guard band of this slab in accumulator units (row scales applied)
    const void* d_slab_image;    // device image of this slab (slab_build)
    uint16_t* S_acc;             // [M] running per-trial hit count, bit 15 = flagged
    int32_t first_slab;
    int32_t row_base;            // first transform row of this slab
    unsigned long long* pair_list;  // (trial << 16 | row): single near-threshold decisions, or nullptr (whole trials get flagged)
    uint32_t* pair_count;
    uint32_t pair_cap;
};
int slab_rows();
int slab_count(int N);
int slab_ncols(int N, int t, bool lower);
size_t slab_image_bytes(int N, bool lower);
void slab_build(int N, bool lower, const double* A, const double* c, uint8_t* image, size_t* offsets, double* row_scale /* N, out */);
cudaError_t slab_launch(const SlabLaunch& l, int num_sms, cudaStream_t s);
cudaError_t slab_count_launch(uint64_t M, const uint16_t* S_acc, int k, int32_t* S_out, uint32_t* flag_list,
                              uint32_t* flag_count, uint32_t flag_cap, uint64_t* block_partials, int grid, cudaStream_t s);

// ---- single-pass tensor-core path for 128 < N <= 1024 (mcr_wide.cu) --------------------------------------
constexpr int WIDE_MAX_SLABS = 8;
struct WideLaunch {
    uint64_t M, path_offset;
    uint32_t key0, key1;
    int32_t N, k, lower;
    const void* d_image;         // wide_build image on the device
    const float* d_c_scaled;     // wide_build thresholds on the device
    double delta_rel[WIDE_MAX_SLABS], delta_abs[WIDE_MAX_SLABS];   // guard band per 128-row slab (relative to ||z||_2), accumulator units
    int32_t* S_out;
    uint32_t* flag_list;
    uint32_t* flag_count;
    uint32_t flag_cap;
    unsigned long long* pair_list;   // trial << 21 | other hits << 10 | row, or nullptr
    uint32_t* pair_count;
    uint32_t pair_cap;
    uint64_t* block_partials;
};
bool wide_supported(int N);
int wide_slabs(int N);
int wide_kblocks(int N);
int wide_trials_per_tile(int N);
size_t wide_image_bytes(int N, bool lower);
void wide_build(int N, bool lower, const double* A, const double* c, uint8_t* image, float* c_scaled, double* row_scale /* N, out */);
int wide_grid(uint64_t M, int N, int num_sms);
cudaError_t wide_launch(const WideLaunch& l, int num_sms, cudaStream_t s);

// ---- streamed tensor-core path: M = trials, normals in TMEM, passes over row ranges (mcr_stream.cu) -----------------
constexpr int STREAM_WMAX = 384;        // transform rows (accumulator columns) per pass
constexpr int STREAM_MAX_PASSES = 4;    // N <= 1024 is at most 3 passes
struct StreamPass { int32_t r0, W, kbn, kend; };   // rows [r0, r0 + W), K-blocks, normal columns [0, kend)
struct StreamLaunch {
    uint64_t M, path_offset;
    uint32_t key0, key1;
    int32_t N, k, lower;
    const void* d_image;         // stream_build image on the device
    int32_t pair;                // CTA-pair form (the image was built for it)
    const double* d_Z;           // supplied normals (M x N fp64, chunk-relative) or nullptr: Philox
    double delta_rel[STREAM_MAX_PASSES], delta_abs[STREAM_MAX_PASSES];   // guard band per pass (relative to ||z||_2), accumulator units
    int32_t* S_out;
    uint32_t* flag_list;
    uint32_t* flag_count;
    uint32_t flag_cap;
    unsigned long long* pair_list;   // trial << 21 | other hits << 10 | row, or nullptr
    uint32_t* pair_count;
    uint32_t pair_cap;
    uint64_t* block_partials;
    // portfolio extension: the image carries one more row (w^T A, w.mu), see stream_build
    int32_t portfolio;
    double port_x, hist_lo, hist_width, port_scale;
    double* port_acc;
    unsigned long long* port_hist;
    double* port_hist_sum;
};
// port = 1: one more image row after the N asset rows (the portfolio return)
bool stream_supported(int N, int port);
int stream_plan(int N, int port, bool lower, StreamPass* pass /* STREAM_MAX_PASSES */);   // number of passes
size_t stream_image_bytes(int N, int port, bool lower);
double stream_mma_flops(int N, int port, bool lower);          // fp16 MMA flops issued per trial
int stream_normals_per_trial(int N, int port, bool lower);     // normals generated per trial (passes re-derive the leading columns)
// port_row: nullptr, or N + 1 values (w^T A, then w.mu).  supplied_order: the image for supplied-normals runs (no Z cache: every pass
// visits its K-blocks in ascending order; generated runs take the cached K-blocks last).  row_scale: out, N (+ 1) entries
// pair: the image of the CTA-pair form (two rank images, each CTA's half of every piece)
void stream_build(int N, bool lower, const double* A, const double* c, const double* port_row, bool supplied_order, bool pair, uint8_t* image, double* row_scale);
bool stream_pair(bool supplied, bool portfolio);   // whether runs of this kind use the CTA-pair form
int stream_grid(uint64_t M, int num_sms, bool pair);
cudaError_t stream_launch(const StreamLaunch& l, int num_sms, cudaStream_t s);

cudaError_t launch_hist(uint64_t M, const int32_t* S, int nbins, unsigned long long* hist, int num_sms, cudaStream_t s);
cudaError_t launch_dump(uint64_t M, uint64_t path_offset, uint32_t key0, uint32_t key1, int N,
                        float* dZ, uint32_t* dU, cudaStream_t s);
cudaError_t launch_fp32_peak(int grid, int iters, float* sink, cudaStream_t s);

}  // namespace mcr
