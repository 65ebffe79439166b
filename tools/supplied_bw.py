#!/usr/bin/env python
"""HBM bandwidth of the supplied-normals mode (M x N fp64 normals resident in HBM, 8 N bytes per trial) per kernel:
    python tools/supplied_bw.py [N ...]          (default 50 100 600)"""
import json
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model

peak = 6557.4
try:
    peak = float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass
e = api.Engine(0)
for N in [int(a) for a in sys.argv[1:]] or [50, 100, 600]:
    m = synthetic_model(N)
    e.upload_sigma(m.mu, m.Sigma)
    M = 20_000_000 if N <= 64 else max(1_000_000, int(1.0e9 // N))
    Z = torch.randn(M, N, dtype=torch.float64, device="cuda")
    ref = None
    for name, k in (("ffma", api.KERNEL_FFMA), ("tensor", api.KERNEL_TENSOR)):
        if k == api.KERNEL_FFMA and N > 128:
            continue
        best = None
        for _ in range(5):
            r = e.run_with_normals_device(Z.data_ptr(), M, m.x, m.k, kernel=k)
            best = r.device_ms if best is None else min(best, r.device_ms)
        ref = ref or r.count
        gbs = M * 8 * N / (best * 1e-3) / 1e9
        print(f"N={N:4d} M={M:9d} {name:6s} ({r.kernel}) {best:8.3f} ms  {gbs:7.1f} GB/s = {gbs / peak:.3f} of {peak:.0f}  count {r.count} {'ok' if r.count == ref else 'MISMATCH'}  rechecked {r.n_rechecked}")
    del Z
    torch.cuda.empty_cache()
