#!/usr/bin/env python
"""Benchmark of the hot path: correlated multi-asset path simulation + tail-event reduction.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): simulations/sec, device-timed.  Headline workload at every N: BASELINE config 2,
"Stoxx50 (50 assets)" shape — synthetic SPD covariance, N=50, M=1e8 trials PER GPU per step
(weak scaling: rank r simulates paths [r*M, (r+1)*M) of one Philox stream, then ONE 64-byte
NCCL all-reduce combines the partial counts; the collective of step n runs on a side stream
while the kernels of step n+1 execute).  A "step" = one full run of M trials per GPU.

One JSON line on rank 0.  `value` = device-timed whole-job sims/s (nothing is HBM-resident:
normals live in registers; the 5 KB transform sits in constant memory).  `e2e` = the same through
the host entry point the C plugin calls (mcr_simulate_host) with HOST buffers: model H2D and
result D2H inside the timed region, wall clock; at N > 1 rank 0 drives all N devices through the
plugin's single-process route (ncclCommInitAll + one grouped all-reduce).  `roofline` is the
roofline of the path kernel: F(N)=N^2+3N algorithmic flops per simulation over its
CUDA-event time, against the FP32 peak (N <= 128) or the measured bf16 tensor peak (N > 128).
`configs` carries the other BASELINE configurations in the same line: C3 (N=100, 1e9 trials in
total, STRONG-scaled over the ranks), C4 (N=600, 1e8 trials in total, tensor-core roofline) and a
reduced C5 sweep.  `cpu_baseline` = the reference's own C-serial / OpenMP models (oracle/_ref,
compiled unmodified against the GSL shim) timed on this box's host cores on a bounded sample;
`gpu_baseline` = the reference's own CUDA model (oracle/_ref/libref_cuda.so: monte_carlos_cuda.cu
compiled unmodified for sm_100a) on the same GPU.
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))
sys.path.insert(0, str(ROOT / "tests"))

N_ASSETS = 50
M_PER_GPU = 100_000_000
SEED = 67


def flops_per_sim(N):
    return N * N + 3 * N  # triangular FMA x2 + mean add + compare


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index=0, period=0.02):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        names = {}
        if nv:
            for n in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown"):
                v = getattr(nv, "nvmlClocksEventReason" + n, None) or getattr(nv, "nvmlClocksThrottleReason" + n, None)
                if v is not None:
                    names[v] = n
        while not self._stop.is_set():
            try:
                if nv:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, n in names.items():
                        if mask & bit:
                            self.reasons.add(n)
                else:
                    import subprocess
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm,"
                                          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [t.strip() for t in out.strip().split(",")]
                    self.samples.append(int(f[0]))
                    self.max_mhz = int(f[1])
                    for n, v in zip(("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap"), f[2:]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=5)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "samples": len(s), "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------ CPU legs
def use_all_host_threads():
    """The OpenMP runtime reads OMP_NUM_THREADS once, when it is loaded (import torch does that), and torchrun exports
    OMP_NUM_THREADS=1 to its workers: set the thread count on the live runtime instead, and report what it says."""
    import ctypes
    ncores = os.cpu_count() or 1
    want = int(os.environ.get("MCR_REF_THREADS", ncores))
    os.environ["OMP_NUM_THREADS"] = str(want)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(want)
        return int(gomp.omp_get_max_threads())
    except OSError:
        return want


def cpu_reference_leg(model, budget_s=12.0):
    """Time the reference's own C-serial and OpenMP models (oracle/_ref) on this box's cores."""
    import oracle_lib
    if not oracle_lib.ref_models_path().exists():
        return None
    ref = oracle_lib.RefModels()
    ncores = use_all_host_threads()
    out = {}
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1), os.dup(2)
    try:
        os.dup2(devnull, 1), os.dup2(devnull, 2)  # the models print progress lines
        for name, fn in (("serial", ref.serial), ("openmp", ref.openmp)):
            M = 100_000
            fn(model.mu, model.Sigma, model.k, model.x, 20_000, SEED, keep_S=False)  # warm
            while True:
                t0 = time.perf_counter()
                r = fn(model.mu, model.Sigma, model.k, model.x, M, SEED, keep_S=False)
                dt = time.perf_counter() - t0
                assert r["rc"] == 0
                if dt > budget_s / 4 or M >= 20_000_000:
                    break
                M = int(M * min(8.0, max(2.0, budget_s / 2 / max(dt, 1e-3))))
            out[name] = {"sims_per_s": M / dt, "M": M, "seconds": dt, "count": r["count"]}
    finally:
        os.dup2(saved[0], 1), os.dup2(saved[1], 2)
        os.close(devnull)
    best = max(out, key=lambda n: out[n]["sims_per_s"])
    return {
        "value": out[best]["sims_per_s"], "unit": "simulations/s",
        "cores": 1 if best == "serial" else ncores, "kind": "reference",
        "sample": f"reference {best} model (oracle/_ref: unmodified monte_carlo_{'serial' if best == 'serial' else 'omp'}.c + GSL-subset shim), "
                  f"N={model.N}, M={out[best]['M']} trials, {out[best]['seconds']:.1f} s wall",
        "serial_sims_per_s": out["serial"]["sims_per_s"],
        "openmp_sims_per_s": out["openmp"]["sims_per_s"], "openmp_threads": ncores,
        "note": "shim timings approximate GSL 2.7; the OpenMP model serialises all RNG draws in one omp critical (omp.c:77-82)",
    }


def gpu_reference_leg(model, our_value):
    """Time the reference's own CUDA model (src/05-GPU/monte_carlos_cuda.cu, compiled unmodified for sm_100a into
    oracle/_ref/libref_cuda.so) on this GPU through its run_model entry point, wall clock as main_runner.c:244-253 does.
    It always allocates, fills and copies back S_values (4 bytes per trial), supports N <= 64, and draws from cuRAND XORWOW."""
    import oracle_lib
    if not oracle_lib.ref_cuda_path().exists():
        return {"unavailable": "oracle/_ref/libref_cuda.so not built (reference tree or nvcc absent at build time)"}
    if model.N > 64:
        return {"unavailable": f"the reference CUDA kernel supports N <= 64 (monte_carlos_cuda.cu:78-81); N={model.N}"}
    ref = oracle_lib.RefCuda()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1), os.dup(2)
    try:
        os.dup2(devnull, 1), os.dup2(devnull, 2)
        ref.cuda(model.mu, model.Sigma, model.k, model.x, 100_000, SEED, keep_S=False)   # context, module load
        M, best = 2_000_000, None
        while True:
            t0 = time.perf_counter()
            r = ref.cuda(model.mu, model.Sigma, model.k, model.x, M, SEED, keep_S=False)
            dt = time.perf_counter() - t0
            assert r["rc"] == 0
            best = (M, dt, r)
            if dt > 2.0 or M >= 100_000_000:
                break
            M = int(min(100_000_000, M * min(8.0, max(2.0, 3.0 / max(dt, 1e-3)))))
        M, dt, r = best
        t0 = time.perf_counter()
        r = ref.cuda(model.mu, model.Sigma, model.k, model.x, M, SEED, keep_S=False)
        dt = min(dt, time.perf_counter() - t0)
    finally:
        os.dup2(saved[0], 1), os.dup2(saved[1], 2)
        os.close(devnull)
    v = M / dt
    return {"value": v, "unit": "simulations/s", "kind": "reference", "M": M, "seconds": dt, "P_hat": r["P_hat"],
            "d2h_bytes": 4 * M + 4, "ratio_ours_device_timed": our_value / v,
            "sample": f"reference CUDA model (monte_carlos_cuda.cu unmodified, nvcc -gencode arch=compute_100a,code=sm_100a, GSL-subset shim), "
                      f"N={model.N}, M={M} trials, wall clock around run_model: cudaMalloc + H2D + kernel + {4 * M / 1e6:.0f} MB S_values D2H + free",
            "note": "one thread per trial, fp64, cuRAND XORWOW with a per-thread curand_init, Z/R in local memory, one-address atomicAdd"}


def run_reference_impl(args):
    from mcr_b200.synthetic import synthetic_model
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    model = synthetic_model(N_ASSETS)
    import oracle_lib
    if not oracle_lib.ref_models_path().exists():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (reference tree was absent at build time)"}))
        return 0
    ref = oracle_lib.RefModels()
    ncores = use_all_host_threads()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1), os.dup(2)
    per = {}
    try:
        os.dup2(devnull, 1), os.dup2(devnull, 2)
        for name, fn, M in (("serial", ref.serial, 400_000), ("openmp", ref.openmp, 200_000)):
            for _ in range(args.warmup):
                fn(model.mu, model.Sigma, model.k, model.x, M // 10, SEED, keep_S=False)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                r = fn(model.mu, model.Sigma, model.k, model.x, M, SEED, keep_S=False)
            dt = time.perf_counter() - t0
            per[name] = {"sims_per_s": M * args.steps / dt, "M": M, "ms_per_step": dt / args.steps * 1e3}
    finally:
        os.dup2(saved[0], 1), os.dup2(saved[1], 2)
        os.close(devnull)
    best = max(per, key=lambda n: per[n]["sims_per_s"])
    v = per[best]["sims_per_s"]
    line = {
        "impl": "reference", "metric": "simulations/sec", "value": v, "unit": "simulations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per[best]["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BASELINE config 2 (Stoxx50 shape): synthetic SPD covariance, N={N_ASSETS}, x=0.02, k=17, seed 67; bounded sample of "
                               f"{per[best]['M']} trials per step on host cores (reference {best} model)"},
        "cpu_baseline": {"value": v, "unit": "simulations/s", "cores": 1 if best == "serial" else ncores,
                         "kind": "reference", "sample": f"{best} model, {per[best]['M']} trials x {args.steps} steps",
                         "serial_sims_per_s": per["serial"]["sims_per_s"], "openmp_sims_per_s": per["openmp"]["sims_per_s"],
                         "openmp_threads": ncores},
        "e2e": {"value": v, "unit": "simulations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------ our arm
def measured_peaks():
    """Driver-written peaks of this pool's B200s (MEASURED_PEAKS.json); B200_PROFILING.md's fallback when absent."""
    mp = ROOT / "MEASURED_PEAKS.json"
    if mp.exists():
        d = json.load(open(mp))
        return {"bf16_burst": d.get("bf16_tflops"), "bf16_sustained": d.get("bf16_tflops_sustained") or d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1590.0, "hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the path kernel, from committed `ncu --set full` captures:
# (kernel, N) -> (bytes, file).  The path kernels read the operand image / constant bank and the code once per launch and write
# two words per CTA, whatever the number of trials (the captures ran 2e7 / 4e6 trials per launch); anything not captured is null.
TRAFFIC_TABLE = {
    ("tensor", 50): (60416, "profiles/r02_tc_path_n50_ncu_summary.txt"),
    ("ffma", 50): (71168, "profiles/r01_ffma_n50_pipelined_ncu_summary.txt"),
    ("tensor", 600): (956416, "profiles/r02_stream_final_n600_ncu_summary.txt"),
}


class Job:
    """One configuration on this rank: a model, its engine state and a pipelined stepper.

    step n = the kernel set of `m_rank` trials on the compute stream, then (world > 1) the 64-byte sum all-reduce of the
    partial words on the communication stream behind an event.  The partial buffers rotate, so step n+1's kernels do not
    wait for step n's collective; `drain` makes the compute stream wait for every outstanding collective."""
    NBUF = 4

    def __init__(self, torch, dist, api, eng, model, A, m_rank, offset, world, stream, comm_stream, kernel=None, sm_reserve=0):
        self.sm_reserve = sm_reserve
        self.torch, self.dist, self.api, self.eng, self.model, self.A = torch, dist, api, eng, model, A
        self.m_rank, self.offset, self.world, self.stream, self.comm = m_rank, offset, world, stream, comm_stream
        self.kernel = api.KERNEL_AUTO if kernel is None else kernel
        self.bufs = [torch.zeros(api.MCR_PARTIAL_WORDS, dtype=torch.int64, device="cuda") for _ in range(self.NBUF)]
        self.done = [torch.cuda.Event() for _ in range(self.NBUF)]
        self.free = [torch.cuda.Event() for _ in range(self.NBUF)]
        self.n = 0

    def upload(self):
        self.eng.upload_model(self.model.mu, self.A, a_is_lower=True)

    def step(self):
        b = self.n % self.NBUF
        self.n += 1
        if self.world > 1:
            self.stream.wait_event(self.free[b])      # the collective that last used this buffer has finished
        self.eng.run_async(self.m_rank, SEED, self.model.x, self.model.k, self.bufs[b].data_ptr(), self.stream.cuda_stream,
                           path_offset=self.offset, kernel=self.kernel, sm_reserve=self.sm_reserve)
        if self.world > 1:
            self.done[b].record(self.stream)
            self.comm.wait_event(self.done[b])
            with self.torch.cuda.stream(self.comm):
                self.dist.all_reduce(self.bufs[b])    # the one cross-GPU exchange: 8 x int64 sums over NCCL
            self.free[b].record(self.comm)
        return b

    def drain(self):
        if self.world > 1:
            for ev in self.free:
                self.stream.wait_event(ev)

    def timed(self, steps, warmup, fence, sampler=None):
        """(ms per step: max over ranks of the CUDA-event time of `steps` steps on the compute stream, collectives
        drained inside the timed region; partial words of the last step)."""
        torch = self.torch
        for _ in range(warmup):
            self.step()
        self.drain()
        fence()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        last = 0
        for _ in range(steps):
            last = self.step()
        self.drain()
        ev1.record(self.stream)
        fence()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            torch.cuda.synchronize()
        return float(t.item()) / steps, self.bufs[last].cpu().numpy().astype(np.uint64)

    def kernel_ms(self, reps=3):
        """(CUDA-event ms of the path kernel per launch, launches per step), events recorded by the library on the stream."""
        out = []
        for _ in range(reps):
            self.step()
            out.append(self.eng.last_main_kernel_ms())
        self.drain()
        self.torch.cuda.synchronize()
        return float(np.mean([m for m, _ in out])), out[0][1]


def tensor_roofline(N, kernel, padded_N, sims_per_launch, k_ms, step_ms, sims_per_step, mma_flops_per_trial, peaks):
    """Roofline of a tensor-core path kernel: algorithmic F(N) flops and executed fp16 MMA flops per trial over the kernel's
    CUDA-event time, against the measured bf16 peak (burst for the kernel alone, sustained for the whole step)."""
    alg = sims_per_launch * flops_per_sim(N) / (k_ms * 1e-3) / 1e12
    alg_step = sims_per_step * flops_per_sim(N) / (step_ms * 1e-3) / 1e12
    tr = TRAFFIC_TABLE.get(("tensor", N))
    r = {"bound": "tensor", "achieved": alg, "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": alg / peaks["bf16_burst"],
         "traffic": tr[0] if tr else None, "traffic_source": tr[1] if tr else None, "kernel": kernel, "kernel_ms_per_launch": k_ms, "flops_per_sim": flops_per_sim(N),
         "frac_of_sustained_peak_over_step": alg_step / peaks["bf16_sustained"], "peak_source": peaks["source"],
         "note": "algorithmic flops F(N)=N^2+3N per simulation; the correlate runs as fp16x3 (hi*hi + lo*hi + hi*lo, fp32 accumulate "
                 "in TMEM), so the tensor pipe executes >= 3x the algorithmic work plus tile padding"}
    if mma_flops_per_trial:
        ex = sims_per_launch * mma_flops_per_trial / (k_ms * 1e-3) / 1e12
        r["executed"] = {"tflops": ex, "frac": ex / peaks["bf16_burst"], "mma_flops_per_trial": mma_flops_per_trial}
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--assets", type=int, default=N_ASSETS, help="asset count of the headline synthetic model (default: BASELINE config 2)")
    ap.add_argument("--paths", type=int, default=M_PER_GPU, help="trials per GPU per step of the headline workload")
    ap.add_argument("--kernel", default="auto", choices=["auto", "ffma", "tensor"], help="path kernel of the headline workload")
    ap.add_argument("--sm-reserve", type=int, default=-1, help="SMs the path kernels leave free for the overlapped all-reduce (default: 1 when N > 1, else 0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3 / C4 / C5 lines (kernel-tuning runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_impl(args)

    import torch
    import torch.distributed as dist
    from mcr_b200 import api
    from mcr_b200.synthetic import synthetic_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        opts = None
        try:   # the collective's kernel must get an SM the moment one frees up at the tail of a path kernel
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            opts = None
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
        cpu_group = dist.new_group(backend="gloo")   # host-side rendezvous that leaves the GPUs idle (e2e leg)

    N, M = args.assets, args.paths
    steps, warmup = args.steps, max(args.warmup, 3)
    peaks = measured_peaks()
    eng = api.Engine(local)
    # explicit (non-default) torch streams: the library launches on `stream`, torch events time it, NCCL runs behind `comm`
    stream = torch.cuda.Stream()
    comm = torch.cuda.Stream(priority=-1)
    torch.cuda.set_stream(stream)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sm_reserve = args.sm_reserve if args.sm_reserve >= 0 else (1 if world > 1 else 0)

    def make_job(n_assets, m_rank, offset, kernel=None):
        mdl = synthetic_model(n_assets)
        A_ = api.build_transform(mdl.Sigma, api.TRANSFORM_LOWER)
        return Job(torch, dist, api, eng, mdl, A_, m_rank, offset, world, stream, comm, kernel, sm_reserve)

    # ---- headline: BASELINE config 2 shape, weak scaling ----
    kernel_sel = {"auto": api.KERNEL_AUTO, "ffma": api.KERNEL_FFMA, "tensor": api.KERNEL_TENSOR}[args.kernel]
    job = make_job(N, M, rank * M, kernel_sel)
    model, A = job.model, job.A
    job.upload()
    with ClockSampler(local) as clocks:
        ms_step, words = job.timed(steps, warmup, fence)
        k_ms, launches_per_kernel_step = job.kernel_ms(min(steps, 5))
    res = api.result_from_partials(words)
    assert res.n_paths == M * world, (res.n_paths, M, world)
    value = M * world / (ms_step * 1e-3)

    # second kernel of the headline shape (BASELINE config 2 names the fp32 FFMA correlate path): same workload, one GPU's share
    alt = None
    if N <= 128 and args.kernel == "auto" and not args.no_configs:
        j2 = make_job(N, M, rank * M, api.KERNEL_FFMA)
        ms2, w2 = j2.timed(max(3, steps // 4), 3, fence)
        k2, l2 = j2.kernel_ms(3)
        r2 = api.result_from_partials(w2)
        assert (r2.count, r2.sum_S) == (res.count, res.sum_S), "FFMA and tensor kernels disagree"
        alt = {"kernel": "ffma_path_kernel (FFMA2 on the FP32 pipe)", "ms_per_step": ms2, "value": M * world / (ms2 * 1e-3),
               "kernel_ms_per_launch": k2, "achieved_tflops": (M / max(1, l2)) * flops_per_sim(N) / (k2 * 1e-3) / 1e12}

    # ---- the other BASELINE configurations, same line (all ranks take part; strong scaling: total work fixed) ----
    configs = []
    if not args.no_configs:
        def run_config(name, n_assets, m_total, c_steps, c_warm, scaling):
            m_rank = m_total // world + (1 if rank < m_total % world else 0)
            off = (m_total // world) * rank + min(rank, m_total % world)
            if scaling == "weak":
                m_rank, off = m_total, rank * m_total
            j = make_job(n_assets, m_rank, off)
            j.upload()
            ms, w = j.timed(c_steps, c_warm, fence)
            km, kl = j.kernel_ms(2)
            r = api.result_from_partials(w)
            total = m_total * world if scaling == "weak" else m_total
            assert r.n_paths == total, (name, r.n_paths, total)
            one = eng.run(min(m_rank, 1 << 20), SEED, j.model.x, j.model.k)
            entry = {"name": name, "N": n_assets, "M_total": total, "scaling": scaling, "n_gpus": world, "steps": c_steps,
                     "warmup": c_warm, "ms_per_step": ms, "value": total / (ms * 1e-3), "unit": "simulations/s",
                     "kernel": one.kernel, "padded_N": one.padded_N, "kernel_ms_per_launch": km, "launches_per_step": kl,
                     "other_ms_per_step": max(0.0, ms - km * kl),
                     "result": {"count": r.count, "sum_S": r.sum_S, "P_hat": r.P_hat, "n_rechecked": r.n_rechecked}}
            sims_per_launch = m_rank / max(1, kl)
            if n_assets > 128:
                entry["roofline"] = tensor_roofline(n_assets, "tc_stream_kernel (tcgen05 fp16x3 correlate, normals in TMEM, DESIGN.md section 3.4)", one.padded_N,
                                                    sims_per_launch, km, ms, m_rank, eng.mma_flops_per_trial(), peaks)
            else:
                tf = sims_per_launch * flops_per_sim(n_assets) / (km * 1e-3) / 1e12
                entry["roofline"] = {"bound": "fp32", "achieved": tf, "unit": "TFLOP/s", "flops_per_sim": flops_per_sim(n_assets)}
            return entry

        configs.append(run_config("C3: EuroNext100 shape, N=100, 1e9 trials in total, strong-scaled", 100, 1_000_000_000, 3, 1, "strong"))
        configs.append(run_config("C4: Stoxx600 shape, N=600, 1e8 trials in total, tensor-core contraction", 600, 100_000_000, 3, 1, "strong"))
        for n_assets in (16, 32, 64, 128, 256, 512, 1024):
            configs.append(run_config(f"C5 (reduced sweep): N={n_assets}, 1e7 trials per GPU", n_assets, 10_000_000, 3, 1, "weak"))
        job.upload()   # back to the headline model for the legs below

    # ---- rank-0-only legs (the other ranks wait on the host, GPUs idle)
    e2e = roof = cpu = gpu_ref = supplied = None
    launches_per_step = 0
    if rank == 0:
        one = eng.run(M, SEED, model.x, model.k, kernel=kernel_sel)
        # fix-up kernels + finalize of one step: device time of this plain single-GPU run of the shard minus its path kernel
        _km, _kl = eng.last_main_kernel_ms()
        fixup_ms = max(0.0, one.device_ms - _km * _kl)
        launches_per_step = one.launches
        # e2e through the host entry point (what the C plugin calls), host buffers, wall clock; N > 1: the plugin's own
        # single-process multi-GPU route over all `world` devices (MCR_NUM_GPUS), rank 0 driving them
        mu_h, A_h = np.ascontiguousarray(model.mu), np.ascontiguousarray(A)
        Mt = M * world
        kw = dict(a_is_lower=True, reupload=True, n_gpus=world, kernel=kernel_sel)
        for _ in range(3):
            api.simulate_host(mu_h, A_h, Mt, SEED, model.x, model.k, **kw)
        nrep = max(3, min(steps, 10))
        t0 = time.perf_counter()
        for _ in range(nrep):
            h = api.simulate_host(mu_h, A_h, Mt, SEED, model.x, model.k, **kw)
        dt = (time.perf_counter() - t0) / nrep
        assert (h.count, h.sum_S) == (res.count, res.sum_S), "plugin route and torchrun route disagree"
        e2e = {"value": Mt / dt, "unit": "simulations/s",
               "h2d_bytes_per_step": int(world * (mu_h.nbytes + 2 * A_h.nbytes + 8 * N * N // 2)),
               "d2h_bytes_per_step": 8 * api.MCR_PARTIAL_WORDS * world, "ms_per_step": dt * 1e3,
               "api": "mcr_simulate_host (host mu/A in, result out; S_values not requested, as the plugin's default)"
                      + (f"; n_gpus={world}: one process drives all devices, ncclCommInitAll + one grouped all-reduce" if world > 1 else "")}
        if world == 1:
            # same call when the caller asks for the reference's per-trial S_values array; once untimed: the first call pays
            # the page faults of a fresh 4*M-byte host array
            api.simulate_host(mu_h, A_h, M, SEED, model.x, model.k, want_S_values=True, **kw)
            t0 = time.perf_counter()
            hs = api.simulate_host(mu_h, A_h, M, SEED, model.x, model.k, want_S_values=True, **kw)
            dt_s = time.perf_counter() - t0
            assert int(hs.S_values.sum(dtype=np.int64)) == one.sum_S
            e2e["with_S_values"] = {"value": M / dt_s, "ms_per_step": dt_s * 1e3, "d2h_bytes_per_step": 4 * M + 64,
                                    "note": "with MonteCarloResult.S_values filled (the reference always does; nothing reads it)"}
        peak_tf = eng.measure_fp32_peak(20.0)
        nominal_tf = 148 * 128 * 2 * (clocks.max_mhz or 1965) * 1e6 / 1e12
        # k_ms is the average over the path-kernel launches of a step (a step of more than one chunk - 2^27 trials, 2^24 on the
        # N > 128 kernels - is several launches): trials per launch = the step's trials / launches
        sims_per_launch = M / max(1, launches_per_kernel_step)
        achieved_tf = sims_per_launch * flops_per_sim(N) / (k_ms * 1e-3) / 1e12
        traffic = TRAFFIC_TABLE.get((one.kernel, N))
        if one.kernel == "tensor" and N > 128:
            roof = tensor_roofline(N, "tc_stream_kernel (tcgen05 fp16x3 correlate, normals in TMEM, DESIGN.md section 3.4)", one.padded_N, sims_per_launch, k_ms,
                                   ms_step, M, eng.mma_flops_per_trial(), peaks)
        else:
            tensor = None
            if one.kernel == "tensor":
                # what the tensor pipe actually executes: 3 passes (fp16 hi/lo split) over the padded dense tile
                exec_tf = sims_per_launch * eng.mma_flops_per_trial() / (k_ms * 1e-3) / 1e12
                tensor = {"executed_tflops": exec_tf, "peak": peaks["bf16_burst"], "frac": exec_tf / peaks["bf16_burst"],
                          "peak_source": peaks["source"],
                          "note": "the correlate runs on tcgen05 (fp16x3, fp32 accumulate in TMEM); the kernel is bound by the Philox/Box-Muller "
                                  "instruction stream, not by the tensor pipe"}
            kname = (f"tc_path_kernel<NOUT={one.padded_N}> (tcgen05 fp16x3 correlate)" if one.kernel == "tensor"
                     else f"ffma_path_kernel<NT={one.padded_N},LOWER,philox>")
            roof = {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved_tf / peak_tf if peak_tf else None,
                    "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None,
                    "kernel": kname, "kernel_ms_per_launch": k_ms, "launches_timed_per_step": launches_per_kernel_step,
                    "flops_per_sim": flops_per_sim(N),
                    "peak_source": "measured here: FFMA2-only kernel (mcr_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 entry",
                    "frac_of_nominal": achieved_tf / nominal_tf, "nominal_peak": nominal_tf, "tensor": tensor, "fp32_ffma_kernel": alt,
                    "note": "FP32-EQUIVALENT: algorithmic flops F(N)=N^2+3N per simulation over the path kernel's CUDA-event time, against "
                            "the FP32 FFMA peak (the metric BASELINE.json names); no per-trial HBM traffic (normals and returns never "
                            "leave the SM).  `fp32_ffma_kernel` is the same workload on the FP32 pipe proper"}
        # supplied-normals leg (north star: "HBM GB/s wherever normals are materialised"): the same correlate / threshold / reduce
        # kernels fed M x N fp64 normals resident in HBM (the layout serial.c:67-69 fills) instead of generating them; 8 N bytes
        # per trial, inputs far larger than L2.  FFMA2 kernel and streamed tcgen05 kernel.
        if world == 1 and not args.no_configs:
            try:
                Mz = 20_000_000 if N <= 64 else max(1_000_000, int(1.0e9 // N))
                Z = torch.randn(Mz, N, dtype=torch.float64, device="cuda")
                supplied = {"M": Mz, "bytes_per_trial": 8 * N, "hbm_peak_gbs": peaks["hbm_gbs"], "peak_source": peaks["source"], "kernels": {}}
                for kname_, ksel in (("ffma", api.KERNEL_FFMA), ("tensor", api.KERNEL_TENSOR)):
                    if ksel == api.KERNEL_FFMA and N > 128:
                        continue
                    best = None
                    for _ in range(4):
                        rz = eng.run_with_normals_device(Z.data_ptr(), Mz, model.x, model.k, kernel=ksel)
                        best = rz.device_ms if best is None else min(best, rz.device_ms)
                    gbs = Mz * 8 * N / (best * 1e-3) / 1e9
                    supplied["kernels"][kname_] = {"device_ms": best, "sims_per_s": Mz / (best * 1e-3), "hbm_gbs": gbs,
                                                   "frac_of_hbm_peak": gbs / peaks["hbm_gbs"], "count": rz.count, "n_rechecked": rz.n_rechecked}
                ks = list(supplied["kernels"].values())
                assert all(k_["count"] == ks[0]["count"] for k_ in ks), "kernels disagree on supplied normals"
                del Z
                torch.cuda.empty_cache()
                # the same mode on a wider model (N = 256, 8 GB of normals): the streamed kernel reads the matrix through TMA boxes
                if N <= 128:
                    Nw, Mw = 256, 3_906_250
                    mw = synthetic_model(Nw)
                    eng.upload_model(mw.mu, api.build_transform(mw.Sigma, api.TRANSFORM_LOWER), a_is_lower=True)
                    Zw = torch.randn(Mw, Nw, dtype=torch.float64, device="cuda")
                    best = None
                    for _ in range(4):
                        rz = eng.run_with_normals_device(Zw.data_ptr(), Mw, mw.x, mw.k)
                        best = rz.device_ms if best is None else min(best, rz.device_ms)
                    gbs = Mw * 8 * Nw / (best * 1e-3) / 1e9
                    supplied["wide"] = {"N": Nw, "M": Mw, "kernel": rz.kernel, "device_ms": best, "sims_per_s": Mw / (best * 1e-3), "hbm_gbs": gbs,
                                        "frac_of_hbm_peak": gbs / peaks["hbm_gbs"], "n_rechecked": rz.n_rechecked}
                    del Zw
                    torch.cuda.empty_cache()
                    job.upload()   # back to the headline model
            except Exception as ex:
                supplied = {"error": repr(ex)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_reference_leg(model)
            except Exception as ex:  # the CPU leg must never take the GPU number down with it
                cpu = {"error": repr(ex)}
            try:
                gpu_ref = gpu_reference_leg(model, value)
            except Exception as ex:
                gpu_ref = {"error": repr(ex)}
    if cpu_group is not None:
        dist.barrier(group=cpu_group)

    if rank == 0:
        kern_total = k_ms * launches_per_kernel_step
        line = {
            "metric": "simulations/sec", "value": value, "unit": "simulations/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if one.kernel == "ffma" else "f16x3+f32acc", "data": "synthetic",
            "config": {"workload": (f"BASELINE config 2 (Stoxx50 shape)" if (N, M) == (N_ASSETS, M_PER_GPU) else f"custom shape (not a BASELINE config)")
                                   + f": synthetic SPD covariance, N={N}, M={M} trials per GPU per step, x={model.x}, k={model.k}, seed {SEED}, "
                                     f"lower-triangular transform, kernel={one.kernel} + fp64 recheck band",
                       "parallelism": (f"paths sharded x{world}; one 64-byte NCCL all-reduce per step on a side stream, overlapped with the "
                                       f"next step's kernels" if world > 1 else "single GPU"),
                       "l2": "no HBM/L2-resident inputs to flush: normals are generated in registers, the transform (a few KB) sits in shared/constant memory"},
            "clocks": clocks.summary(),
            "kernel_ms_per_step": kern_total, "comm_ms_per_step": max(0.0, ms_step - kern_total - fixup_ms), "fixup_ms_per_step": fixup_ms,
            "sm_reserve": sm_reserve,
            "e2e": e2e, "gpu_launches": launches_per_step * steps,
            "roofline": roof, "cpu_baseline": cpu, "gpu_baseline": gpu_ref, "supplied_normals": supplied, "configs": configs,
            # SURVEY.md section 8(d) secondary lines: normals consumed per second (N per trial) and the dense-equivalent flop rate
            "secondary": {"normals_per_s": value * N, "dense_equivalent_tflops": value * (2 * N * N + 2 * N) / 1e12,
                          "note": "whole job, over the step time; N normals and 2N^2+2N flops per simulation"},
            "result": {"count": res.count, "sum_S": res.sum_S, "P_hat": res.P_hat, "n_rechecked": res.n_rechecked},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
