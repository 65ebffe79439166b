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
#include <sys/mman.h>

#include <algorithm>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
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

// ---- S_values on demand: compact on the device, streamed, widened on the host ---------------
// The reference's models return S (an int per trial, src/model_interface.h:12; monte_carlos_cuda.cu:138-139 copies M ints back
// in one piece after the kernel).  Here a run that wants S_values is enqueued in parts of S_PART trials: the path kernels write
// the part's S into an int32 scratch, a small kernel narrows it to one byte per trial (two when N > 255) into one of two
// compact buffers, a copy stream moves that into pinned host memory while the next part is being simulated, and host threads
// widen it into the caller's int array.  PCIe carries 1 byte per trial instead of 4, nothing is copied into pageable memory by
// the driver, and only the last part's copy and widening are not hidden behind the kernels.
constexpr uint64_t S_PART = 1ull << 23;

template <typename T>
__global__ void __launch_bounds__(256) narrow_S_kernel(const int32_t* __restrict__ in, T* __restrict__ out, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            const int4 v = *reinterpret_cast<const int4*>(in + i);
            if (sizeof(T) == 1) *reinterpret_cast<uchar4*>(out + i) = make_uchar4((unsigned char)v.x, (unsigned char)v.y, (unsigned char)v.z, (unsigned char)v.w);
            else *reinterpret_cast<ushort4*>(out + i) = make_ushort4((unsigned short)v.x, (unsigned short)v.y, (unsigned short)v.z, (unsigned short)v.w);
        } else {
            for (uint64_t j = i; j < n; ++j) out[j] = (T)in[j];
        }
    }
}

static int ensure_S_stream(mcr_engine* e, uint64_t part) {
    if (e->S_part_cap >= part) return 0;
    cudaFree(e->d_S_part); e->d_S_part = nullptr;
    for (int b = 0; b < 2; ++b) {
        cudaFree(e->d_S_compact[b]); e->d_S_compact[b] = nullptr;
        if (e->h_S_compact[b]) cudaFreeHost(e->h_S_compact[b]);
        e->h_S_compact[b] = nullptr;
    }
    e->S_part_cap = 0;
    MCR_CUDA(cudaMalloc(&e->d_S_part, sizeof(int32_t) * part));
    for (int b = 0; b < 2; ++b) {
        MCR_CUDA(cudaMalloc(&e->d_S_compact[b], 2 * part));
        MCR_CUDA(cudaHostAlloc(&e->h_S_compact[b], 2 * part, cudaHostAllocPortable));
        if (!e->ev_compact[b]) MCR_CUDA(cudaEventCreateWithFlags(&e->ev_compact[b], cudaEventDisableTiming));
        if (!e->ev_copied[b]) MCR_CUDA(cudaEventCreateWithFlags(&e->ev_copied[b], cudaEventDisableTiming));
    }
    if (!e->copy_stream) MCR_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    e->S_part_cap = part;
    return 0;
}

// compact host buffer -> the caller's int array, on a few host threads (the destination is pageable memory the size of 4 M bytes)
template <typename T>
static void widen_S(const T* src, int32_t* dst, uint64_t n) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nt = (unsigned)std::min<uint64_t>(std::min(8u, hw), std::max<uint64_t>(1, n >> 18));
    auto work = [=](uint64_t a, uint64_t b) { for (uint64_t i = a; i < b; ++i) dst[i] = (int32_t)src[i]; };
    if (nt <= 1) { work(0, n); return; }
    std::vector<std::thread> th;
    const uint64_t per = (n + nt - 1) / nt;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, std::min(n, t * per), std::min(n, (t + 1) * per));
    work(0, std::min(n, per));
    for (auto& t : th) t.join();
}

}  // namespace mcr

using namespace mcr;

// One host thread per device, kept for the life of the process (a call used to start and join its threads twice: ~0.2 ms of a
// 5.6 ms run on 8 GPUs).  Worker g sleeps on a condition variable until run() hands it `job(g)`.  The pool is never destroyed
// (no static destructor runs while workers wait); calls are serialised by g_host_mu.
namespace {
class DeviceWorkers {
public:
    void run(int n, const std::function<void(int)>& job) {
        std::unique_lock<std::mutex> lk(mu_);
        while ((int)threads_ < n) {
            const int g = threads_++;
            std::thread([this, g]() { loop(g); }).detach();
        }
        job_ = &job;
        pending_ = n;
        ++generation_;
        want_ = n;
        cv_work_.notify_all();
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int g) {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_work_.wait(lk, [&] { return generation_ != seen && g < want_; });
            seen = generation_;
            const std::function<void(int)>* job = job_;
            lk.unlock();
            (*job)(g);
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)>* job_ = nullptr;
    uint64_t generation_ = 0;
    int threads_ = 0, pending_ = 0, want_ = 0;
};
DeviceWorkers& device_workers() {
    static DeviceWorkers* w = new DeviceWorkers();   // leaked on purpose
    return *w;
}
}  // namespace

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
    std::vector<uint64_t> m(n), off(n);
    uint64_t acc = 0;
    for (int g = 0; g < n; ++g) {
        m[g] = req->M / n + ((uint64_t)g < req->M % n ? 1 : 0);
        off[g] = acc;
        acc += m[g];
    }
    int rc = 0;
    const bool want_ext = req->opt.want_hist || req->opt.want_portfolio;
    const bool narrow8 = req->N <= 255;   // S <= N fits a byte

    for (int g = 0; g < n && !rc; ++g)
        if (get_engine(req->first_device + g, &eng[g])) { rc = -1; break; }
    // per device: upload the model when it changed (or on request), then whatever `work` enqueues.  With several devices every
    // device gets its own host thread: the host side of a shard - validity checks, transposes, operand image, a handful of
    // synchronous copies, ~60 us - would otherwise run n times in sequence before the last device even starts (8 GPUs: 0.8 ms of a
    // 5.6 ms run).  Error messages are thread-local, so a worker's message is carried back and re-raised here.
    auto on_every_device = [&](const std::function<int(int)>& work) -> int {
        if (n == 1) return work(0);
        std::vector<int> rcs(n, 0);
        std::vector<std::string> msgs(n);
        device_workers().run(n, [&](int g) {
            rcs[g] = work(g);
            if (rcs[g]) msgs[g] = mcr_last_error();
        });
        for (int g = 0; g < n; ++g)
            if (rcs[g]) return fail("%s", msgs[g].c_str());
        return 0;
    };
    if (!rc)
        rc = on_every_device([&](int g) -> int {
            if (req->reupload || !same_model(eng[g], req)) return mcr_upload_model(eng[g], req->N, req->mu, req->A, req->a_is_lower);
            return 0;
        });
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
    auto enqueue_ext = [&](mcr_engine* e) -> int {
        // extension partials of this shard in the all-reducible layout of mcr_run_async_ext
        if (!e->d_ext_counts && (cudaMalloc(&e->d_ext_counts, sizeof(uint64_t) * MCR_EXT_COUNT_WORDS) != cudaSuccess ||
                                 cudaMalloc(&e->d_ext_sums, sizeof(double) * MCR_EXT_SUM_WORDS) != cudaSuccess))
            return fail("extension partial buffers: allocation failed on device %d", e->device);
        return enqueue_gather_ext(e, e->d_ext_counts, e->d_ext_sums, e->stream);
    };
    for (int g = 0; g < n && !rc; ++g) {
        cudaSetDevice(eng[g]->device);
        cudaEventRecord(eng[g]->ev0, eng[g]->stream);
        eng[g]->accumulate = false;
    }
    if (!req->want_S_values) {
        if (!rc)
            rc = on_every_device([&](int g) -> int {
                mcr_engine* e = eng[g];
                cudaSetDevice(e->device);
                mcr_run_options o = req->opt;
                o.struct_size = sizeof(o);
                o.d_S_values = nullptr;
                int r = enqueue_run(e, m[g], off[g], req->seed, req->x, req->k, &o, nullptr, e->d_partials, e->stream);
                if (!r && want_ext) r = enqueue_ext(e);
                return r;
            });
    } else {
        // The caller's array is usually fresh from malloc (the reference allocates it per run): M ints = 100 000 first-touch page
        // faults per 1e8 trials for the widening threads.  Ask for transparent huge pages on its 2 MB-aligned interior (a hint: a no-op
        // where THP is off or the pages are already there).
        {
            static const bool no_thp = [] { const char* v = getenv("MCR_NO_HUGEPAGE"); return v && *v && atoi(v) != 0; }();
            const uintptr_t lo = (reinterpret_cast<uintptr_t>(req->S_values_out) + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
            const uintptr_t hi = (reinterpret_cast<uintptr_t>(req->S_values_out) + sizeof(int32_t) * req->M) & ~(uintptr_t)((2u << 20) - 1);
            if (!no_thp && hi > lo) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
        }
        // S_values: the shard of every device in parts of S_PART trials; part j is narrowed and copied out while part j + 1 runs,
        // and widened into the caller's array by this thread (and helpers) while the GPUs work on
        uint64_t parts = 0;
        for (int g = 0; g < n; ++g) parts = std::max<uint64_t>(parts, (m[g] + S_PART - 1) / S_PART);
        for (int g = 0; g < n && !rc; ++g) {
            cudaSetDevice(eng[g]->device);
            rc = ensure_S_stream(eng[g], S_PART);
        }
        auto widen_part = [&](uint64_t j) -> int {   // wait for part j of every device and widen it
            for (int g = 0; g < n; ++g) {
                const uint64_t a = j * S_PART;
                if (a >= m[g]) continue;
                const uint64_t cnt = std::min<uint64_t>(S_PART, m[g] - a);
                mcr_engine* e = eng[g];
                cudaSetDevice(e->device);
                if (cudaEventSynchronize(e->ev_copied[j & 1]) != cudaSuccess) return fail("device %d: S_values copy failed: %s", e->device, cudaGetErrorString(cudaGetLastError()));
                if (narrow8) widen_S(reinterpret_cast<const uint8_t*>(e->h_S_compact[j & 1]), req->S_values_out + off[g] + a, cnt);
                else widen_S(reinterpret_cast<const uint16_t*>(e->h_S_compact[j & 1]), req->S_values_out + off[g] + a, cnt);
            }
            return 0;
        };
        for (uint64_t j = 0; j < std::max<uint64_t>(parts, 1) && !rc; ++j) {
            for (int g = 0; g < n && !rc; ++g) {
                mcr_engine* e = eng[g];
                const uint64_t a = j * S_PART;
                if (a >= m[g] && !(j == 0)) continue;   // (an empty shard still runs once, so that its partial words are zeroed)
                const uint64_t cnt = a < m[g] ? std::min<uint64_t>(S_PART, m[g] - a) : 0;
                cudaSetDevice(e->device);
                mcr_run_options o = req->opt;
                o.struct_size = sizeof(o);
                o.d_S_values = e->d_S_part;
                rc = enqueue_run(e, cnt, off[g] + a, req->seed, req->x, req->k, &o, nullptr, e->d_partials, e->stream);
                e->accumulate = true;
                if (rc || cnt == 0) continue;
                const int b = (int)(j & 1);
                if (j >= 2) cudaStreamWaitEvent(e->stream, e->ev_copied[b], 0);   // the compact buffer's previous content has left the device
                const int grid = (int)std::min<uint64_t>((uint64_t)e->num_sms * 8, (cnt / 4 + 255) / 256 + 1);
                if (narrow8) narrow_S_kernel<uint8_t><<<grid, 256, 0, e->stream>>>(e->d_S_part, e->d_S_compact[b], cnt);
                else narrow_S_kernel<uint16_t><<<grid, 256, 0, e->stream>>>(e->d_S_part, reinterpret_cast<uint16_t*>(e->d_S_compact[b]), cnt);
                cudaEventRecord(e->ev_compact[b], e->stream);
                cudaStreamWaitEvent(e->copy_stream, e->ev_compact[b], 0);
                cudaMemcpyAsync(e->h_S_compact[b], e->d_S_compact[b], cnt * (narrow8 ? 1 : 2), cudaMemcpyDeviceToHost, e->copy_stream);
                cudaEventRecord(e->ev_copied[b], e->copy_stream);
                ++e->last_launches;
                if (cudaGetLastError() != cudaSuccess) rc = fail("device %d: S_values streaming failed", e->device);
            }
            if (!rc && j >= 1) rc = widen_part(j - 1);
        }
        for (int g = 0; g < n && !rc; ++g) {
            eng[g]->accumulate = false;
            cudaSetDevice(eng[g]->device);
            if (want_ext) rc = enqueue_ext(eng[g]);
        }
        if (!rc && parts >= 1) rc = widen_part(parts - 1);
    }
    for (int g = 0; g < n; ++g) if (eng[g]) eng[g]->accumulate = false;
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
    return rc;
}
