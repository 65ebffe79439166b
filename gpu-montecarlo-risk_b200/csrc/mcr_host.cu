// One-shot host entry point: what the get_cuda_model plugin (src/06-B200/monte_carlo_b200.c)
// calls in place of the reference's cuda_simulate (src/05-GPU/monte_carlos_cuda.cu:75-163).
//
// Differences from the reference wrapper, by design:
//  - engines (context, stream, scratch) are created once per device and cached for the process,
//    so a second run_model call pays no cudaMalloc / context cost;
//  - S_values (M ints, 400 MB at M = 1e8) is produced only when asked for;
//  - the path range can be sharded over n_gpus devices of this process; the per-device partial
//    words are combined with ONE NCCL all-reduce (8 x uint64, latency-bound over NVLink); the extension outputs
//    (histogram of S, portfolio sums and histogram), when requested, with one more per buffer (uint64 / fp64).
//    NCCL is resolved with dlopen at first multi-GPU use, so single-GPU users need no libnccl.
#include "mcr_engine.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace mcr {

// ---- minimal NCCL binding (types as in nccl.h 2.x) ------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclUint64 = 5;
constexpr int kNcclFloat64 = 8;
constexpr int kNcclSum = 0;
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_host_mu;
static std::map<int, mcr_engine*> g_engines;                 // device -> cached engine
static std::map<std::pair<int, int>, std::vector<ncclComm_t>> g_comms;  // (first_device, n) -> comms

static int load_nccl() {
    if (g_nccl.lib) return 0;
    const char* names[] = {getenv("MCR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm) continue;
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return fail("multi-GPU run needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
#define MCR_SYM(field, name)                                                   \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                        \
    if (!g_nccl.field) return fail("NCCL symbol %s not found", name);
    MCR_SYM(CommInitAll, "ncclCommInitAll")
    MCR_SYM(AllReduce, "ncclAllReduce")
    MCR_SYM(GroupStart, "ncclGroupStart")
    MCR_SYM(GroupEnd, "ncclGroupEnd")
    MCR_SYM(GetErrorString, "ncclGetErrorString")
#undef MCR_SYM
    return 0;
}

static int get_engine(int device, mcr_engine** out) {
    auto it = g_engines.find(device);
    if (it != g_engines.end()) { *out = it->second; return 0; }
    mcr_engine* e = nullptr;
    if (mcr_create(device, &e)) return -1;
    g_engines[device] = e;
    *out = e;
    return 0;
}

static bool same_model(const mcr_engine* e, const mcr_host_request* r) {
    return e->N == r->N && e->lower == (r->a_is_lower != 0) &&
           memcmp(e->mu.data(), r->mu, sizeof(double) * r->N) == 0 &&
           memcmp(e->A.data(), r->A, sizeof(double) * (size_t)r->N * r->N) == 0;
}

}  // namespace mcr

using namespace mcr;

extern "C" int mcr_simulate_host(const mcr_host_request* req, mcr_result* out) {
    if (!req || !out) return fail("mcr_simulate_host: null argument");
    if (req->struct_size != sizeof(mcr_host_request)) return fail("mcr_simulate_host: struct_size mismatch (ABI)");
    if (!req->mu || !req->A || req->N <= 0) return fail("mcr_simulate_host: bad model");
    if (req->want_S_values && !req->S_values_out) return fail("mcr_simulate_host: want_S_values without a buffer");
    std::lock_guard<std::mutex> lk(g_host_mu);
    const int n = std::max(1, req->n_gpus);
    int have = 0;
    if (mcr_device_count(&have)) return -1;
    if (req->first_device < 0 || req->first_device + n > have)
        return fail("mcr_simulate_host: asked for devices %d..%d but %d are visible", req->first_device, req->first_device + n - 1, have);

    std::vector<mcr_engine*> eng(n, nullptr);
    std::vector<int32_t*> dS(n, nullptr);
    std::vector<uint64_t> m(n), off(n);
    uint64_t acc = 0;
    for (int g = 0; g < n; ++g) {
        m[g] = req->M / n + ((uint64_t)g < req->M % n ? 1 : 0);
        off[g] = acc;
        acc += m[g];
    }
    int rc = 0;
    const bool want_ext = req->opt.want_hist || req->opt.want_portfolio;
    auto cleanup = [&]() { for (int g = 0; g < n; ++g) if (dS[g]) { cudaSetDevice(eng[g]->device); cudaFree(dS[g]); } };

    for (int g = 0; g < n && !rc; ++g) {
        if (get_engine(req->first_device + g, &eng[g])) { rc = -1; break; }
        if ((req->reupload || !same_model(eng[g], req)) && mcr_upload_model(eng[g], req->N, req->mu, req->A, req->a_is_lower)) { rc = -1; break; }
    }
    if (n > 1 && !rc) {
        if (load_nccl()) rc = -1;
        else if (!g_comms.count({req->first_device, n})) {
            std::vector<int> devs(n);
            for (int g = 0; g < n; ++g) devs[g] = req->first_device + g;
            std::vector<ncclComm_t> comms(n);
            ncclResult_t r = g_nccl.CommInitAll(comms.data(), n, devs.data());
            if (r != 0) rc = fail("ncclCommInitAll failed: %s", g_nccl.GetErrorString(r));
            else g_comms[{req->first_device, n}] = comms;
        }
    }
    // enqueue every shard, then the all-reduce, then wait
    for (int g = 0; g < n && !rc; ++g) {
        mcr_engine* e = eng[g];
        cudaSetDevice(e->device);
        mcr_run_options o = req->opt;
        o.struct_size = sizeof(o);
        o.d_S_values = nullptr;
        if (req->want_S_values && m[g]) {
            if (cudaMalloc(&dS[g], sizeof(int32_t) * m[g]) != cudaSuccess) { rc = fail("S_values allocation (%llu ints) failed on device %d", (unsigned long long)m[g], e->device); break; }
            o.d_S_values = dS[g];
        }
        cudaEventRecord(e->ev0, e->stream);
        rc = enqueue_run(e, m[g], off[g], req->seed, req->x, req->k, &o, nullptr, e->d_partials, e->stream);
        if (!rc && want_ext) {
            // extension partials of this shard in the all-reducible layout of mcr_run_async_ext
            if (!e->d_ext_counts && (cudaMalloc(&e->d_ext_counts, sizeof(uint64_t) * MCR_EXT_COUNT_WORDS) != cudaSuccess ||
                                     cudaMalloc(&e->d_ext_sums, sizeof(double) * MCR_EXT_SUM_WORDS) != cudaSuccess)) {
                rc = fail("extension partial buffers: allocation failed on device %d", e->device);
                break;
            }
            rc = enqueue_gather_ext(e, e->d_ext_counts, e->d_ext_sums, e->stream);
        }
    }
    if (n > 1 && !rc) {
        std::vector<ncclComm_t>& comms = g_comms[{req->first_device, n}];
        g_nccl.GroupStart();
        for (int g = 0; g < n; ++g) {
            ncclResult_t r = g_nccl.AllReduce(eng[g]->d_partials, eng[g]->d_partials, MCR_PARTIAL_WORDS, kNcclUint64, kNcclSum, comms[g], eng[g]->stream);
            if (r != 0 && !rc) rc = fail("ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
            if (want_ext) {
                r = g_nccl.AllReduce(eng[g]->d_ext_counts, eng[g]->d_ext_counts, MCR_EXT_COUNT_WORDS, kNcclUint64, kNcclSum, comms[g], eng[g]->stream);
                if (r != 0 && !rc) rc = fail("ncclAllReduce (extension counts) failed: %s", g_nccl.GetErrorString(r));
                r = g_nccl.AllReduce(eng[g]->d_ext_sums, eng[g]->d_ext_sums, MCR_EXT_SUM_WORDS, kNcclFloat64, kNcclSum, comms[g], eng[g]->stream);
                if (r != 0 && !rc) rc = fail("ncclAllReduce (extension sums) failed: %s", g_nccl.GetErrorString(r));
            }
        }
        ncclResult_t r = g_nccl.GroupEnd();
        if (r != 0 && !rc) rc = fail("ncclGroupEnd failed: %s", g_nccl.GetErrorString(r));
    }
    float ms_max = 0;
    int launches = 0;
    for (int g = 0; g < n && !rc; ++g) {
        mcr_engine* e = eng[g];
        cudaSetDevice(e->device);
        cudaEventRecord(e->ev1, e->stream);
        cudaMemcpyAsync(e->h_partials, e->d_partials, sizeof(uint64_t) * MCR_PARTIAL_WORDS, cudaMemcpyDeviceToHost, e->stream);
        if (dS[g]) cudaMemcpyAsync(req->S_values_out + off[g], dS[g], sizeof(int32_t) * m[g], cudaMemcpyDeviceToHost, e->stream);
        if (want_ext && g == 0) {   // after the all-reduce every device holds the totals
            e->h_ext_counts.resize(MCR_EXT_COUNT_WORDS);
            e->h_ext_sums.resize(MCR_EXT_SUM_WORDS);
            cudaMemcpyAsync(e->h_ext_counts.data(), e->d_ext_counts, sizeof(uint64_t) * MCR_EXT_COUNT_WORDS, cudaMemcpyDeviceToHost, e->stream);
            cudaMemcpyAsync(e->h_ext_sums.data(), e->d_ext_sums, sizeof(double) * MCR_EXT_SUM_WORDS, cudaMemcpyDeviceToHost, e->stream);
        }
    }
    for (int g = 0; g < n && !rc; ++g) {
        mcr_engine* e = eng[g];
        cudaSetDevice(e->device);
        cudaError_t err = cudaStreamSynchronize(e->stream);
        if (err != cudaSuccess) { rc = fail("device %d: %s", e->device, cudaGetErrorString(err)); break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e->ev0, e->ev1);
        ms_max = std::max(ms_max, ms);
        launches += e->last_launches;
    }
    if (!rc) {
        rc = want_ext ? mcr_result_from_partials_ext(eng[0]->h_partials, eng[0]->h_ext_counts.data(), eng[0]->h_ext_sums.data(),
                                                     eng[0]->port_hist_lo, eng[0]->port_hist_hi, out)
                      : mcr_result_from_partials(eng[0]->h_partials, out);
        out->device_ms = ms_max;
        out->kernel_used = eng[0]->last_kernel;
        out->padded_N = eng[0]->last_NT;
        out->launches = launches;
        if (!rc && out->n_paths != req->M) rc = fail("internal: reduced %llu paths, expected %llu", (unsigned long long)out->n_paths, (unsigned long long)req->M);
    }
    cleanup();
    return rc;
}
