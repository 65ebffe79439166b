// Tensor-core kernels, host side of the N <= 128 kernel (variant table; the kernel itself is in
// mcr_tc_path.cuh, one instantiation per mcr_tc_inst.cu translation unit) and the row-slab kernel for N > 128.
#include "mcr_tc_path.cuh"

namespace mcr {

// ======================================================================================================
// Large asset counts (N > 128): row slabs.
//
// 128 trials x N accumulators do not fit TMEM (512 columns) and the transform does not fit shared memory,
// so the correlate is done one slab of 256 transform rows at a time: slab t needs the normals of
// columns [0, ncols_t) only (ncols_t = min(N, 256 (t+1)) for a lower-triangular transform), which the
// generator simply re-derives from the counter-based stream, and its 256 x 64 transform K-blocks are
// streamed from L2 into the same ring slot as the Z K-block they multiply.  Each launch adds the slab's
// hit count into a per-trial uint16 (bit 15 = "inside the guard band in some slab"); slab_count_kernel
// then reduces that array exactly like the single-launch kernels do.  Costs: the normals are generated
// about (slabs + 1) / 2 times, and 4 bytes of HBM traffic per trial and slab.
// ======================================================================================================
struct TcSlabParams {
    uint64_t M, path_offset;
    uint32_t key0, key1;
    int32_t ncols;        // normal columns this slab multiplies; the constant 1 sits in column ncols
    int32_t kb_count;     // K-blocks = ceil((ncols + 1) / 64)
    float delta_rel, delta_abs;   // scaled guard band = delta_rel * ||z||_2 + delta_abs (Cauchy-Schwarz bound on sum |a||z|)
    float thr_max;                // upper bound of that band over all trials: every |z_j| < 7 (Box-Muller on 32-bit uniforms)
    const uint4* B_slab;  // global: kb_count x [hi 32 KB | lo 32 KB], SLAB_ROWS rows x 128 B, pre-swizzled
    uint16_t* S_acc;      // [M] running hit count, bit 15 = flagged
    int32_t first_slab;   // overwrite S_acc instead of accumulating
    int32_t row_base;     // first transform row of this slab
    unsigned long long* pair_list;   // (trial << 16 | row) of single near-threshold decisions, fixed up per pair
    uint32_t* pair_count;
    uint32_t pair_cap;
};

constexpr int SLAB_ROWS = 256;                            // transform rows per slab = UMMA N = TMEM columns of one accumulator region
constexpr uint32_t SLAB_Z_BYTES = 2 * tc::TILE_M * 128;    // Z K-block: hi + lo
constexpr uint32_t SLAB_B_BYTES = 2 * SLAB_ROWS * 128;     // B K-block: hi + lo
constexpr uint32_t SLAB_SLOT_BYTES = SLAB_Z_BYTES + SLAB_B_BYTES;
constexpr int SLAB_KSPLIT = 4;                             // generator threads per trial row (each owns 16 of a K-block's 64 columns)
constexpr int SLAB_GEN_THREADS = 128 * SLAB_KSPLIT;
constexpr int SLAB_THREADS = SLAB_GEN_THREADS + 32;
constexpr uint32_t SLAB_KEY_MASK = 0xFFFFFC00u;            // epilogue keys: bits 10.. = |accumulator| << 1 (truncated)

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory"); }

// 16 generator warps: warp w works on trial rows 32 (w % 4) .. +31 (its TMEM lane quadrant) and on columns
// 16 (w / 4) .. +15 of every K-block while generating, accumulator columns 64 (w / 4) .. +63 in the epilogue.
__global__ void __launch_bounds__(SLAB_THREADS, 1) tc_slab_kernel(const TcSlabParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t s_empty[2], s_done[2];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_part_S[2][SLAB_KSPLIT][128], s_part_k1[2][SLAB_KSPLIT][128], s_part_k2[2][SLAB_KSPLIT][128];
    __shared__ float s_part_zmax[2][SLAB_KSPLIT][128];
    // warp index through a shuffle: warp-uniform for the compiler, so the MMA warp keeps its descriptors in uniform registers
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    constexpr uint32_t TMEM_COLS = 2 * SLAB_ROWS;   // two accumulator regions of SLAB_ROWS columns: all of TMEM
    constexpr int MMA_WARP = SLAB_GEN_THREADS / 32;
    constexpr int EPI_COLS = SLAB_ROWS / SLAB_KSPLIT;   // accumulator columns per epilogue thread

    if (tid == 0)
        for (int b = 0; b < 2; ++b) { mbar_init(&s_empty[b], 1); mbar_init(&s_done[b], 1); }
    if (warp == MMA_WARP) tmem_alloc(&s_tmem, TMEM_COLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    const uint64_t n_tiles = (p.M + TILE_M - 1) / TILE_M;
    const uint64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    constexpr uint32_t IDESC = make_idesc(SLAB_ROWS);
    const int KBN = p.kb_count;

    if (warp == MMA_WARP) {
        for (uint64_t n = 0; n < my_tiles; ++n) {
            for (int kb = 0; kb < KBN; ++kb) {
                const uint32_t u = (uint32_t)(n * KBN + kb), slot = u & 1;
                named_bar_sync(1 + slot, SLAB_THREADS);
                if (elect_one()) {
                    tc_fence_after();
                    // descriptors once per K-block; the 16-column K-steps add 2 (32 bytes >> 4) to the address field
                    const uint64_t zhi_d = make_desc(smem_u32(smem + slot * SLAB_SLOT_BYTES)), zlo_d = zhi_d + (uint64_t)(TILE_M * 128 / 16);
                    const uint64_t bhi_d = zhi_d + (uint64_t)(SLAB_Z_BYTES / 16), blo_d = bhi_d + (uint64_t)(SLAB_ROWS * 128 / 16);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(n & 1) * SLAB_ROWS;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint64_t za = pass == 2 ? zlo_d : zhi_d;
                        const uint64_t ba = pass == 1 ? blo_d : bhi_d;
#pragma unroll
                        for (int ks = 0; ks < KBLOCK / 16; ++ks)
                            umma_f16(d_tmem, za + 2 * ks, ba + 2 * ks, IDESC, (kb | pass | ks) != 0 ? 1u : 0u);
                    }
                    umma_commit(&s_empty[slot]);
                    if (kb == KBN - 1) umma_commit(&s_done[n & 1]);
                }
                __syncwarp();
            }
        }
    } else {
        const int row = tid & 127, part = tid >> 7;   // part: K-split index while generating, column quarter in the epilogue
        const uint32_t ring_row = smem_u32(smem) + (uint32_t)(row / 8) * 1024u + (uint32_t)(row % 8) * 128u;
        const uint32_t r7s = (uint32_t)(row % 8) << 4;
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)part * EPI_COLS;
        const int n_full = p.ncols / 8;
        uint32_t empty_phase = 0, done_phase = 0;
        float prev_zmax = 0.0f;
        uint64_t prev_i = 0;

        auto epilogue = [&](uint64_t n, uint64_t i, float zmax) {
            const uint32_t r = (uint32_t)(n & 1);
            mbar_wait(&s_done[r], (done_phase >> r) & 1u);
            done_phase ^= 1u << r;
            tc_fence_after();
            // Per accumulator a sortable key: |v| (14 mantissa bits) in the high bits, the column in bits 1..8, the sign
            // in bit 0 (rotate left by one, then overwrite the low mantissa bits).  The two smallest keys of the
            // trial's row tell whether none, exactly one (which one, and how the fast path decided it) or several
            // of its decisions are inside the guard band.
            uint32_t S = 0, k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
#pragma unroll
            for (int c0 = 0; c0 < EPI_COLS; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + r * SLAB_ROWS + c0, v);
                tmem_ld_wait();
                float m = 3.0e38f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    S += v[c] >> 31;
                    m = fminf(m, fabsf(__uint_as_float(v[c])));
                }
                // keys only where something can be inside the band (thr_max bounds every trial's band from above)
                if (m < p.thr_max) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const uint32_t key = (__funnelshift_l(v[c], v[c], 1) & (SLAB_KEY_MASK | 1u)) | ((uint32_t)(c0 + c) << 1);
                        k2 = min(k2, max(k1, key));
                        k1 = min(k1, key);
                    }
                }
            }
            tc_fence_before();
            const uint32_t part_cols = (uint32_t)(part * EPI_COLS) << 1;
            s_part_S[r][part][row] = S;
            s_part_k1[r][part][row] = k1 == 0xFFFFFFFFu ? k1 : k1 + part_cols;   // column bits: local column + first column of this quarter (no carry: < 256)
            s_part_k2[r][part][row] = k2 == 0xFFFFFFFFu ? k2 : k2 + part_cols;
            s_part_zmax[r][part][row] = zmax;
            named_bar_sync(3, SLAB_GEN_THREADS);   // the four column quarters of every row are in
            if (part == 0 && i < p.M) {
                uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;
                S = 0; zmax = 0.0f;
#pragma unroll
                for (int q = 0; q < SLAB_KSPLIT; ++q) {
                    S += s_part_S[r][q][row];
                    zmax += s_part_zmax[r][q][row];   // total sum of squares of the trial's normals in this slab
                    const uint32_t a = s_part_k1[r][q][row], b = s_part_k2[r][q][row];
                    m2 = min(m2, max(m1, a)); m1 = min(m1, a);
                    m2 = min(m2, max(m1, b)); m1 = min(m1, b);
                }
                const float thr = __fmaf_rn(p.delta_rel, sqrtf(zmax) * 1.0001f, p.delta_abs);
                const uint32_t thr_key = __float_as_uint(thr) << 1;
                uint32_t flagged = 0u;
                if ((m1 & SLAB_KEY_MASK) < thr_key) {
                    if ((m2 & SLAB_KEY_MASK) < thr_key || !p.pair_list) {
                        flagged = 0x8000u;           // several near-threshold decisions: the whole trial goes to the fp64 kernel
                    } else {
                        S -= m1 & 1u;                // leave the one uncertain decision out; exact_pairs_kernel adds the fp64 one
                        const uint32_t slot = atomicAdd(p.pair_count, 1u);
                        if (slot < p.pair_cap)
                            p.pair_list[slot] = ((unsigned long long)i << 16) | (unsigned long long)(p.row_base + ((m1 >> 1) & 0xFFu));
                    }
                }
                const uint32_t old = p.first_slab ? 0u : (uint32_t)p.S_acc[i];
                p.S_acc[i] = (uint16_t)(((old & 0x7FFFu) + S) | (old & 0x8000u) | flagged);
            }
        };

        for (uint64_t n = 0; n < my_tiles; ++n) {
            const uint64_t tile = blockIdx.x + (uint64_t)gridDim.x * n;
            const uint64_t i = tile * TILE_M + row;
            const uint64_t path = p.path_offset + i;
            PhiloxPath rng;
            rng.init((uint32_t)path, (uint32_t)(path >> 32), p.key1);
            float zmax = 0.0f;
#pragma unroll 1
            for (int kb = 0; kb < KBN; ++kb) {
                const uint32_t u = (uint32_t)(n * KBN + kb), slot = u & 1;
                if (u >= 2) {
                    mbar_wait(&s_empty[slot], (empty_phase >> slot) & 1u);
                    empty_phase ^= 1u << slot;
                }
                // the slab's transform K-block: 64 KB from the pre-swizzled global image (L2 resident), asynchronous
                {
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.B_slab) + (size_t)kb * SLAB_B_BYTES;
                    const uint32_t dst = smem_u32(smem + slot * SLAB_SLOT_BYTES + SLAB_Z_BYTES);
#pragma unroll
                    for (int q = 0; q < (int)(SLAB_B_BYTES / 16 / SLAB_GEN_THREADS); ++q)
                        cp_async16(dst + 16 * (tid + SLAB_GEN_THREADS * q), src + 16 * (tid + SLAB_GEN_THREADS * q));
                }
                const uint32_t zhi_row = ring_row + slot * SLAB_SLOT_BYTES, zlo_row = zhi_row + TILE_M * 128;
                // this thread's two chunks (16 columns) of the K-block
                const int c0 = 8 * kb + 2 * part;
                if (c0 + 2 <= n_full) {
                    float z[16];
                    rng.normals_multi<4>((uint32_t)(2 * c0), p.key0, p.key1, z);
                    float unused = 0.0f;
#pragma unroll
                    for (int q = 0; q < 16; ++q) zmax = fmaf(z[q], z[q], zmax);   // here "zmax" carries sum z^2
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        float zc[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) zc[q] = z[8 * w + q];
                        store_chunk(zc, 2 * part + w, zhi_row, zlo_row, r7s, unused);
                    }
                } else {
                    for (int c8 = c0; c8 < c0 + 2; ++c8) {   // chunks touching the end of the normal columns
                        float z[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) z[q] = 0.0f;
                        if (8 * c8 < p.ncols) rng.normals((uint32_t)(2 * c8), p.key0, p.key1, z);
                        if (8 * c8 + 4 < p.ncols) rng.normals((uint32_t)(2 * c8 + 1), p.key0, p.key1, z + 4);
                        float zmax_unused = 0.0f;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (8 * c8 + q < p.ncols) zmax = fmaf(z[q], z[q], zmax);
                            if (8 * c8 + q == p.ncols) z[q] = 1.0f;
                            if (8 * c8 + q > p.ncols) z[q] = 0.0f;
                        }
                        store_chunk(z, c8 - 8 * kb, zhi_row, zlo_row, r7s, zmax_unused);
                    }
                }
                cp_async_wait_all();
                fence_async_smem();
                if (kb == 0) tc_fence_before();
                named_bar_arrive(1 + slot, SLAB_THREADS);
            }
            if (n > 0) epilogue(n - 1, prev_i, prev_zmax);
            prev_i = i;
            prev_zmax = zmax;
        }
        if (my_tiles > 0) epilogue(my_tiles - 1, prev_i, prev_zmax);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// reduce the per-trial slab sums: flagged trials go to the fp64 fix-up list, the rest are counted
__global__ void __launch_bounds__(256) slab_count_kernel(uint64_t M, const uint16_t* S_acc, int k, int32_t* S_out,
                                                         uint32_t* flag_list, uint32_t* flag_count, uint32_t flag_cap,
                                                         uint64_t* block_partials) {
    uint32_t cnt = 0, sum_s = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = S_acc[i];
        if (v & 0x8000u) {
            const uint32_t slot = atomicAdd(flag_count, 1u);
            if (slot < flag_cap) flag_list[slot] = (uint32_t)i;
        } else {
            cnt += (v >= (uint32_t)k) ? 1u : 0u;
            sum_s += v;
            if (S_out) S_out[i] = (int32_t)v;
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    sum_s = __reduce_add_sync(0xffffffffu, sum_s);
    __shared__ uint32_t s_c[8], s_s[8];
    if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = cnt; s_s[threadIdx.x >> 5] = sum_s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t c = 0, s = 0;
        for (int w = 0; w < 8; ++w) { c += s_c[w]; s += s_s[w]; }
        block_partials[2 * (uint64_t)blockIdx.x] = c;
        block_partials[2 * (uint64_t)blockIdx.x + 1] = s;
    }
}

static const TcVariant* tc_supplied() {
    static const TcVariant v = tc_variant_inst<TC_SUPPLIED_INST>();
    return &v;
}

// smallest instantiation with nout >= N and 64 * kb >= N + 1
static const TcVariant* tc_find(int N) {
    static const TcVariant table[TC_VARIANTS] = {
        tc_variant_inst<0>(), tc_variant_inst<1>(), tc_variant_inst<2>(), tc_variant_inst<3>(), tc_variant_inst<4>(),
        tc_variant_inst<5>(), tc_variant_inst<6>(), tc_variant_inst<7>(), tc_variant_inst<8>(), tc_variant_inst<9>(),
    };
    for (const TcVariant& v : table)
        if (v.nout >= N && 64 * v.kb >= N + 1) return &v;
    return nullptr;
}

int tc_max_n() { return 128; }

bool tc_supported(int rows) { return tc_find(rows) != nullptr; }

// ---- the supplied-normals instantiation (one shape for every N <= 128) ----
size_t tc_supplied_image_bytes() { return tc_supplied()->image_bytes; }
int tc_supplied_padded_n() { return tc_supplied()->nout; }
void tc_supplied_build(int N, const double* A, const double* c, uint8_t* image, double* row_scale) { tc_supplied()->build(A, c, N, nullptr, image, row_scale); }
cudaError_t tc_supplied_occupancy(int* bps) { return tc_supplied()->occupancy(bps, false); }

size_t tc_image_bytes(int rows) { const TcVariant* v = tc_find(rows); return v ? v->image_bytes : 0; }

int tc_padded_n(int rows) { const TcVariant* v = tc_find(rows); return v ? v->nout : 0; }

void tc_build(int N, int rows, const double* A, const double* c, const double* port_row, uint8_t* image, double* row_scale) {
    tc_find(rows)->build(A, c, N, port_row, image, row_scale);
}
double tc_z_factor() { return TC_Z_FACTOR; }

cudaError_t tc_occupancy(int rows, bool portfolio, int* bps) { return tc_find(rows)->occupancy(bps, portfolio); }

cudaError_t tc_launch(int N, int rows, const TcLaunch& l, int grid, cudaStream_t s) {
    TcParams p{};
    p.M = l.M; p.path_offset = l.path_offset; p.key0 = l.key0; p.key1 = l.key1; p.k = l.k; p.N = N;
    p.keys.set(l.key0, l.key1);
    // the band arrives in accumulator units (row scales applied); with MCR_TC_FOLD the kernel's bound on |z| is in units of sqrt(2 ln 2)
    p.delta_rel = (float)(l.delta_rel * TC_Z_FACTOR * (TC_FOLD ? 1.0 + 1e-6 : 1.0)); p.delta_abs = (float)l.delta_abs;
    p.B_image = reinterpret_cast<const uint4*>(l.d_image); p.S_out = l.S_out;
    p.flag_list = l.flag_list; p.flag_count = l.flag_count; p.flag_cap = l.flag_cap; p.block_partials = l.block_partials;
    p.pair_list = l.pair_list; p.pair_count = l.pair_count; p.pair_cap = l.pair_cap;
    p.portfolio = l.portfolio; p.port_thresh = (float)(-l.port_x * l.port_scale); p.port_inv_scale = (float)(1.0 / l.port_scale);
    p.hist_lo = (float)l.hist_lo; p.hist_inv_w = (float)(1.0 / l.hist_width); p.hist_lo_d = l.hist_lo; p.hist_w_d = l.hist_width;
    p.port_acc = l.port_acc; p.port_hist = l.port_hist; p.port_hist_sum = l.port_hist_sum;
    for (int q = 0; q < 8; ++q) {
        const int col = 8 * (N / 8) + q;
        p.tail_mul[q] = col < N ? 1.0f : 0.0f;
        p.tail_add[q] = col == N ? 1.0f : 0.0f;
    }
    p.Z = l.d_Z;
    return (l.d_Z ? tc_supplied() : tc_find(rows))->launch(p, grid, s);
}

// ---- slab path host side ---------------------------------------------------------------------------------
int slab_rows() { return SLAB_ROWS; }
int slab_count(int N) { return (N + SLAB_ROWS - 1) / SLAB_ROWS; }
int slab_ncols(int N, int t, bool lower) { return lower ? std::min(N, SLAB_ROWS * (t + 1)) : N; }
int slab_kblocks(int ncols) { return (ncols + 1 + tc::KBLOCK - 1) / tc::KBLOCK; }

size_t slab_image_bytes(int N, bool lower) {
    size_t b = 0;
    for (int t = 0; t < slab_count(N); ++t) b += (size_t)slab_kblocks(slab_ncols(N, t, lower)) * SLAB_B_BYTES;
    return b;
}

// image of every slab, back to back; offsets[t] = byte offset of slab t (offsets has slab_count(N) + 1 entries)
void slab_build(int N, bool lower, const double* A, const double* c, uint8_t* image, size_t* offsets, double* row_scale) {
    for (int i = 0; i < N; ++i) {
        double mx = std::fabs(c[i]);
        for (int k = 0; k < N; ++k) mx = std::max(mx, std::fabs(A[(size_t)i * N + k]));
        row_scale[i] = tc_row_scale(mx);
    }
    size_t off = 0;
    for (int t = 0; t < slab_count(N); ++t) {
        offsets[t] = off;
        const int ncols = slab_ncols(N, t, lower), kbn = slab_kblocks(ncols);
        uint8_t* img = image + off;
        memset(img, 0, (size_t)kbn * SLAB_B_BYTES);
        for (int r = 0; r < SLAB_ROWS; ++r) {
            const int i = SLAB_ROWS * t + r;
            for (int k = 0; k <= ncols; ++k) {
                double v;
                if (k < ncols) v = (i < N) ? A[(size_t)i * N + k] : 0.0;
                else v = (i < N) ? c[i] : TC_PAD_ROW;
                if (i < N) v *= row_scale[i];
                const __half hi = __float2half_rn((float)v);
                const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
                // K-block kb = k / 64: [hi: 128 rows x 128 B][lo: ...], same swizzle as the device-side Z rows
                const int kb = k / tc::KBLOCK, kk = k % tc::KBLOCK;
                const size_t o = (size_t)kb * SLAB_B_BYTES + (size_t)(r / 8) * 1024 + (size_t)(r % 8) * 128 +
                                 (size_t)(((kk / 8) ^ (r % 8)) * 16) + (size_t)(kk % 8) * 2;
                memcpy(img + o, &hi, 2);
                memcpy(img + o + SLAB_ROWS * 128, &lo, 2);
            }
        }
        off += (size_t)kbn * SLAB_B_BYTES;
    }
    offsets[slab_count(N)] = off;
}

cudaError_t slab_launch(const SlabLaunch& l, int num_sms, cudaStream_t s) {
    const size_t smem = 2 * (size_t)SLAB_SLOT_BYTES + 1024;
    cudaError_t e = cudaFuncSetAttribute(tc_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    TcSlabParams p{};
    p.M = l.M; p.path_offset = l.path_offset; p.key0 = l.key0; p.key1 = l.key1;
    p.ncols = l.ncols; p.kb_count = slab_kblocks(l.ncols);
    p.delta_rel = (float)l.delta_rel; p.delta_abs = (float)l.delta_abs;   // already in accumulator units
    p.thr_max = p.delta_rel * 7.0f * sqrtf((float)l.ncols) * 1.001f + p.delta_abs;
    p.B_slab = reinterpret_cast<const uint4*>(l.d_slab_image); p.S_acc = l.S_acc; p.first_slab = l.first_slab;
    p.row_base = l.row_base; p.pair_list = l.pair_list; p.pair_count = l.pair_count; p.pair_cap = l.pair_cap;
    const uint64_t tiles = (l.M + tc::TILE_M - 1) / tc::TILE_M;
    const int grid = (int)std::min<uint64_t>((uint64_t)num_sms, tiles ? tiles : 1);
    tc_slab_kernel<<<grid, SLAB_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t slab_count_launch(uint64_t M, const uint16_t* S_acc, int k, int32_t* S_out, uint32_t* flag_list,
                              uint32_t* flag_count, uint32_t flag_cap, uint64_t* block_partials, int grid, cudaStream_t s) {
    slab_count_kernel<<<grid, 256, 0, s>>>(M, S_acc, k, S_out, flag_list, flag_count, flag_cap, block_partials);
    return cudaGetLastError();
}

}  // namespace mcr
