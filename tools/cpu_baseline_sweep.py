#!/usr/bin/env python
"""The reference's own C-serial and OpenMP models (oracle/_ref, unmodified sources + GSL-subset shim) timed on this
box's host cores across asset counts — the CPU side of BASELINE config 5.  Wall clock like main_runner.c:244-253."""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200")); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib
from mcr_b200.synthetic import synthetic_model
ncores = os.cpu_count() or 1
os.environ.setdefault("OMP_NUM_THREADS", str(ncores))
ref = oracle_lib.RefModels()
devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1), os.dup(2)
rows = []
for N in (16, 32, 50, 64, 100, 128, 256, 600, 1024):
    m = synthetic_model(N)
    M = max(2000, int(3e9 / (N * N + 60 * N)) // 10)
    out = {"N": N, "M": M, "cores": ncores}
    os.dup2(devnull, 1); os.dup2(devnull, 2)
    try:
        for name, fn in (("serial", ref.serial), ("openmp", ref.openmp)):
            fn(m.mu, m.Sigma, m.k, m.x, M // 10, 67, keep_S=False)
            t0 = time.perf_counter(); r = fn(m.mu, m.Sigma, m.k, m.x, M, 67, keep_S=False); dt = time.perf_counter() - t0
            out[name + "_sims_per_s"] = M / dt
    finally:
        os.dup2(saved[0], 1); os.dup2(saved[1], 2)
    rows.append(out); print(json.dumps(out), flush=True)
