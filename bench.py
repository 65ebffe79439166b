#!/usr/bin/env python
"""Benchmark of the hot path: correlated multi-asset path simulation + tail-event reduction.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): simulations/sec, device-timed.  Workload at every N: BASELINE config 2,
"Stoxx50 (50 assets)" shape — synthetic SPD covariance, N=50, M=1e8 trials PER GPU per step
(weak scaling: rank r simulates paths [r*M, (r+1)*M) of one Philox stream, then ONE 64-byte
NCCL all-reduce combines the partial counts).  A "step" = one full run of M trials per GPU.

One JSON line on rank 0.  `value` = device-timed whole-job sims/s (nothing is HBM-resident:
normals live in registers; the 5 KB transform sits in constant memory).  `e2e` = the same through
the host entry point the C plugin calls (mcr_simulate_host) with HOST buffers: model H2D and
result D2H inside the timed region, wall clock.  `roofline` is the FP32-pipe roofline of the path
kernel: F(N)=N^2+3N algorithmic flops per simulation over its CUDA-event time.
`cpu_baseline` = the reference's own C-serial / OpenMP models (oracle/_ref, compiled unmodified
against the GSL shim) timed on this box's host cores on a bounded sample.
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
        "config": {"workload": f"Stoxx50-shape synthetic SPD, N={N_ASSETS}, x=0.02, k=17, seed 67; bounded sample of "
                               f"{per[best]['M']} trials per step on host cores"},
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
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--assets", type=int, default=N_ASSETS, help="asset count of the synthetic model (default: BASELINE config 2)")
    ap.add_argument("--paths", type=int, default=M_PER_GPU, help="trials per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    N, M = args.assets, args.paths
    model = synthetic_model(N)
    eng = api.Engine(local)
    A = eng.upload_sigma(model.mu, model.Sigma, api.TRANSFORM_LOWER)
    partials = torch.zeros(api.MCR_PARTIAL_WORDS, dtype=torch.int64, device="cuda")
    # an explicit (non-default) torch stream: the library launches on it, torch events time it, NCCL uses it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def step():
        eng.run_async(M, SEED, model.x, model.k, partials.data_ptr(), stream.cuda_stream, path_offset=rank * M)
        if world > 1:
            dist.all_reduce(partials)  # the one cross-GPU exchange: 8 x int64 sums over NCCL

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    fence()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        fence()
        ms_total = ev0.elapsed_time(ev1)
        # per-launch CUDA-event time of the path kernel (events recorded by the library on the same stream)
        for _ in range(min(args.steps, 5)):
            step()
            kernel_ms.append(eng.last_main_kernel_ms())
        fence()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    words = partials.cpu().numpy().astype(np.uint64)
    res = api.result_from_partials(words)
    assert res.n_paths == M * world, (res.n_paths, M, world)
    ms_step = ms_total / args.steps
    value = M * world / (ms_step * 1e-3)
    launches_per_step = None

    # ---- rank-0-only legs
    e2e = None
    roof = None
    cpu = None
    if rank == 0:
        one = eng.run(M, SEED, model.x, model.k)
        launches_per_step = one.launches
        k_ms = float(np.mean([m for m, n in kernel_ms]))
        launches_per_kernel_step = kernel_ms[0][1]
        # e2e through the host entry point (what the C plugin calls), host buffers, wall clock
        if world == 1:
            mu_h, A_h = np.ascontiguousarray(model.mu), np.ascontiguousarray(A)
            for _ in range(3):
                api.simulate_host(mu_h, A_h, M, SEED, model.x, model.k, a_is_lower=True, reupload=True)
            nrep = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(nrep):
                h = api.simulate_host(mu_h, A_h, M, SEED, model.x, model.k, a_is_lower=True, reupload=True)
            dt = (time.perf_counter() - t0) / nrep
            assert (h.count, h.sum_S) == (one.count, one.sum_S)
            # same call when the caller asks for the reference's per-trial S_values array (4*M bytes D2H); once untimed:
            # the first call pays the page faults of a fresh 4*M-byte host array
            api.simulate_host(mu_h, A_h, M, SEED, model.x, model.k, a_is_lower=True, reupload=True, want_S_values=True)
            t0 = time.perf_counter()
            hs = api.simulate_host(mu_h, A_h, M, SEED, model.x, model.k, a_is_lower=True, reupload=True, want_S_values=True)
            dt_s = time.perf_counter() - t0
            assert int(hs.S_values.sum()) == one.sum_S
            e2e_s = {"value": M / dt_s, "ms_per_step": dt_s * 1e3, "d2h_bytes_per_step": 4 * M + 64,
                     "note": "with MonteCarloResult.S_values filled (the reference always does; nothing reads it)"}
            e2e = {"value": M / dt, "unit": "simulations/s", "with_S_values": e2e_s,
                   "h2d_bytes_per_step": int(mu_h.nbytes + A_h.nbytes + 8 * N * N // 2),
                   "d2h_bytes_per_step": 8 * api.MCR_PARTIAL_WORDS, "ms_per_step": dt * 1e3,
                   "api": "mcr_simulate_host (host mu/A in, result out; S_values not requested, as the plugin's default)"}
        peak_tf = eng.measure_fp32_peak(20.0)
        nominal_tf = 148 * 128 * 2 * (clocks.max_mhz or 1965) * 1e6 / 1e12
        # k_ms is the average over the path-kernel launches of a step (a step of more than one chunk - 2^27 trials, 2^24 on the
        # N > 128 kernels - is several launches): trials per launch = the step's trials / launches
        sims_per_launch = M / max(1, launches_per_kernel_step)
        achieved_tf = sims_per_launch * flops_per_sim(N) / (k_ms * 1e-3) / 1e12
        tensor = None
        if one.kernel == "tensor" and N <= 128:
            # what the tensor pipe actually executes: 3 passes (fp16 hi/lo split) over the padded dense tile
            kpad = 64 * ((N + 1 + 63) // 64)
            exec_tf = sims_per_launch * 3 * 2 * one.padded_N * kpad / (k_ms * 1e-3) / 1e12
            mp = ROOT / "MEASURED_PEAKS.json"
            bf16_peak = json.load(open(mp)).get("bf16_tflops") if mp.exists() else 1590.0
            tensor = {"executed_tflops": exec_tf, "peak": bf16_peak, "frac": exec_tf / bf16_peak,
                      "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst, of measured)" if mp.exists() else "fallback 1.59 PFLOP/s (of fallback)",
                      "note": "the correlate runs on tcgen05 (fp16x3, fp32 accumulate in TMEM); the kernel is bound by the Philox/Box-Muller "
                              "instruction stream, not by the tensor pipe"}
        kname = ("tc_wide_kernel / tc_slab_kernel (tcgen05 fp16x3 correlate, N > 128: see DESIGN.md section 3)" if one.kernel == "tensor" and N > 128
                 else f"tc_path_kernel<NOUT={one.padded_N}> (tcgen05 fp16x3 correlate)" if one.kernel == "tensor"
                 else f"ffma_path_kernel<NT={one.padded_N},LOWER,philox>")
        roof = {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf if peak_tf else None,
                "traffic": 61440 if one.kernel == "tensor" else 71168,  # dram__bytes_read+write per launch, ncu --set full (profiles/r01_bench_tc_n50_ncu_summary.txt)
                "kernel": kname, "kernel_ms_per_launch": k_ms,
                "launches_timed_per_step": launches_per_kernel_step,
                "flops_per_sim": flops_per_sim(N),
                "peak_source": "measured here: FFMA2-only kernel (mcr_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 entry",
                "frac_of_nominal": achieved_tf / nominal_tf, "nominal_peak": nominal_tf, "tensor": tensor,
                "note": "algorithmic flops F(N)=N^2+3N per simulation over the path kernel's CUDA-event time, against the FP32 FFMA peak "
                        "(the metric BASELINE.json names); no per-trial HBM traffic (normals and returns never leave the SM)"}
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_reference_leg(model)
            except Exception as ex:  # the CPU leg must never take the GPU number down with it
                cpu = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": "simulations/sec", "value": value, "unit": "simulations/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if (launches_per_step and one.kernel == "ffma") else "f16x3+f32acc", "data": "synthetic",
            "config": {"workload": f"Stoxx50-shape synthetic SPD (BASELINE config 2), N={N}, M={M} trials per GPU per step, "
                                   f"x={model.x}, k={model.k}, seed {SEED}, lower-triangular transform, kernel={one.kernel if rank == 0 else '?'} + fp64 recheck band",
                       "parallelism": f"paths sharded x{world}, one 64-byte NCCL all-reduce per step" if world > 1 else "single GPU",
                       "l2": "no HBM/L2-resident inputs to flush: normals are generated in registers, the transform (a few KB) sits in shared/constant memory"},
            "clocks": clocks.summary(),
            "e2e": e2e, "gpu_launches": (launches_per_step or 0) * args.steps,
            "roofline": roof, "cpu_baseline": cpu,
            "result": {"count": res.count, "sum_S": res.sum_S, "P_hat": res.P_hat, "n_rechecked": res.n_rechecked},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
