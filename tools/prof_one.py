#!/usr/bin/env python
"""Run one kernel flavour a few times (for ncu):  python tools/prof_one.py N kernel M [reps]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model
N, kern, M = int(sys.argv[1]), {"ffma": api.KERNEL_FFMA, "tensor": api.KERNEL_TENSOR, "exact": api.KERNEL_EXACT, "auto": api.KERNEL_AUTO}[sys.argv[2]], int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
m = synthetic_model(N)
e = api.Engine(0)
e.upload_sigma(m.mu, m.Sigma)
for _ in range(reps):
    r = e.run(M, 67, m.x, m.k, kernel=kern)
print(r.kernel, r.padded_N, r.device_ms, e.last_main_kernel_ms(), r.count, r.n_rechecked)
