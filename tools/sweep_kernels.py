#!/usr/bin/env python
"""FFMA vs tensor-core kernel across asset counts (kernel-only CUDA-event time, whole run time)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model
e = api.Engine(0)
Ns = [int(a) for a in sys.argv[1:]] or [8, 16, 24, 32, 40, 48, 50, 64, 80, 100, 128]
print(f"{'N':>4s} {'kernel':>7s} {'NT':>4s} {'kernel ms':>10s} {'run ms':>9s} {'sims/s':>11s} {'normals/s':>11s} {'TF-equiv':>9s} {'rechecked':>9s}")
for N in Ns:
    m = synthetic_model(N); e.upload_sigma(m.mu, m.Sigma)
    M = 40_000_000 if N <= 64 else 20_000_000
    for name, kern in (("ffma", api.KERNEL_FFMA), ("tensor", api.KERNEL_TENSOR)):
        for _ in range(2): e.run(M, 67, m.x, m.k, kernel=kern)
        ks, rs = [], []
        for _ in range(5):
            r = e.run(M, 67, m.x, m.k, kernel=kern); ks.append(e.last_main_kernel_ms()[0]); rs.append(r.device_ms)
        k_ms, r_ms = float(np.median(ks)), float(np.median(rs))
        print(f"{N:4d} {name:>7s} {r.padded_N:4d} {k_ms:10.3f} {r_ms:9.3f} {M / r_ms * 1e3:11.3e} {M * N / r_ms * 1e3:11.3e} {M * (N * N + 3 * N) / r_ms * 1e3 / 1e12:9.2f} {r.n_rechecked:9d}", flush=True)
