#!/usr/bin/env python
"""One small run of the streamed kernel (and the tile kernel) for compute-sanitizer racecheck / synccheck."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model
e = api.Engine(0)
for N, M in ((600, 700), (50, 2000)):
    m = synthetic_model(N)
    e.upload_sigma(m.mu, m.Sigma)
    r = e.run(M, 67, m.x, m.k)
    x = e.run(M, 67, m.x, m.k, kernel=api.KERNEL_EXACT)
    assert (r.count, r.sum_S) == (x.count, x.sum_S)
    print(N, r.kernel, r.count, "ok")
# supplied normals through the TMA boxes (N > 128, even N) and through element loads (odd N)
import numpy as np
for N, M in ((600, 500), (301, 300)):
    m = synthetic_model(N)
    e.upload_sigma(m.mu, m.Sigma)
    Z = np.random.default_rng(N).standard_normal((M, N))
    r = e.run_with_normals(Z, m.x, m.k, kernel=api.KERNEL_TENSOR)
    x = e.run_with_normals(Z, m.x, m.k, kernel=api.KERNEL_EXACT)
    assert (r.count, r.sum_S) == (x.count, x.sum_S) and (r.S_values == x.S_values).all()
    print(N, "supplied", r.kernel, r.count, "ok")
e.close()
