#!/usr/bin/env python
"""A few supplied-normals runs of one kernel for ncu:  python tools/prof_supplied.py N [M] [auto|ffma|tensor]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model
N = int(sys.argv[1])
M = int(sys.argv[2]) if len(sys.argv) > 2 else max(1_000_000, int(1.0e9 // N))
kern = {"auto": api.KERNEL_AUTO, "ffma": api.KERNEL_FFMA, "tensor": api.KERNEL_TENSOR}[sys.argv[3] if len(sys.argv) > 3 else "tensor"]
m = synthetic_model(N)
e = api.Engine(0)
e.upload_sigma(m.mu, m.Sigma)
Z = torch.randn(M, N, dtype=torch.float64, device="cuda")
for _ in range(3):
    r = e.run_with_normals_device(Z.data_ptr(), M, m.x, m.k, kernel=kern)
print(N, M, r.kernel, r.device_ms, "ms", M * 8 * N / r.device_ms / 1e6, "GB/s")
