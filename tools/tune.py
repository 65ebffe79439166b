#!/usr/bin/env python
"""Kernel-tuning helper: time the path kernel of every development build under devlibs/*/ (see
scratch/devbuild.sh) on the same workload and check that all of them return identical counts.
    python tools/tune.py [N] [M]            # parent: one subprocess per library
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))


def child(N, M):
    import numpy as np
    from mcr_b200 import api
    from mcr_b200.synthetic import synthetic_model
    m = synthetic_model(N)
    e = api.Engine(0)
    e.upload_sigma(m.mu, m.Sigma)
    nr = bool(os.environ.get("MCR_TUNE_NORECHECK"))   # ablated kernels produce garbage decisions
    kern = {"auto": api.KERNEL_AUTO, "ffma": api.KERNEL_FFMA, "tensor": api.KERNEL_TENSOR}[os.environ.get("MCR_TUNE_KERNEL", "auto")]
    for _ in range(3):
        r = e.run(M, 67, m.x, m.k, no_recheck=nr, kernel=kern)
    ks, tot = [], []
    for _ in range(7):
        r = e.run(M, 67, m.x, m.k, no_recheck=nr, kernel=kern)
        avg, launches = e.last_main_kernel_ms()   # average per launch; a run is split into chunks
        ks.append(avg * launches)
        tot.append(r.device_ms)
    raw = e.run(M, 67, m.x, m.k, no_recheck=True, kernel=kern)
    kms = float(np.median(ks))
    F = N * N + 3 * N
    print(json.dumps({"kernel_ms": kms, "run_ms": float(np.median(tot)), "norecheck_run_ms": raw.device_ms,
                      "sims_per_s": M / kms * 1e3, "tflops": M * F / kms * 1e3 / 1e12,
                      "count": r.count, "sum_S": r.sum_S, "rechecked": r.n_rechecked, "NT": r.padded_N}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
    libs = sorted((ROOT / "devlibs").glob("*/libmcr_b200.so"))
    if os.environ.get("MCR_TUNE_LIBS"):   # comma-separated subset of devlibs/<name>
        keep = set(os.environ["MCR_TUNE_LIBS"].split(","))
        libs = [l for l in libs if l.parent.name in keep]
    ref = None
    for lib in libs:
        env = dict(os.environ, MCR_B200_LIB=str(lib))
        try:
            p = subprocess.run([sys.executable, __file__, "--child", str(N), str(M)], env=env, capture_output=True, text=True,
                               timeout=float(os.environ.get("MCR_TUNE_CHILD_TIMEOUT", "120")))
        except subprocess.TimeoutExpired:
            print(f"{lib.parent.name:12s} TIMEOUT")
            continue
        if p.returncode != 0:
            print(f"{lib.parent.name:12s} FAILED: {p.stderr.strip().splitlines()[-1] if p.stderr.strip() else p.returncode}")
            continue
        d = json.loads(p.stdout.strip().splitlines()[-1])
        ref = ref or (d["count"], d["sum_S"])
        ok = "ok" if (d["count"], d["sum_S"]) == ref else "MISMATCH"
        print(f"{lib.parent.name:12s} kernel {d['kernel_ms']:8.3f} ms  run {d['run_ms']:8.3f} ms  norecheck {d['norecheck_run_ms']:8.3f} ms  "
              f"{d['sims_per_s']:.3e} sims/s  {d['tflops']:6.2f} TF  rechecked {d['rechecked']}  {ok}")
