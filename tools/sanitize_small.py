#!/usr/bin/env python
"""Small runs of every fast kernel for compute-sanitizer (memcheck): python tools/sanitize_small.py"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model

e = api.Engine(0)
for N, M in ((15, 3000), (50, 3000), (128, 1500), (200, 1500), (600, 1100), (1024, 600)):
    m = synthetic_model(N)
    A = e.upload_sigma(m.mu, m.Sigma)
    r = e.run(M, 67, m.x, m.k)
    x = e.run(M, 67, m.x, m.k, kernel=api.KERNEL_EXACT)
    assert (r.count, r.sum_S) == (x.count, x.sum_S), (N, r, x)
    h = api.simulate_host(m.mu, A, M, 67, m.x, m.k, a_is_lower=True, want_S_values=True, want_hist=True)
    assert (h.count, h.sum_S) == (x.count, x.sum_S)
    if N >= 36:
        p = e.run(M, 67, m.x, m.k, portfolio_threshold=0.01)
        assert (p.count, p.sum_S) == (x.count, x.sum_S) and int(p.portfolio["hist"].sum()) == M
    Z = np.random.default_rng(1).standard_normal((300, N))
    a = e.run_with_normals(Z, m.x, m.k, kernel=api.KERNEL_TENSOR)
    b = e.run_with_normals(Z, m.x, m.k, kernel=api.KERNEL_EXACT)
    assert (a.count, a.sum_S) == (b.count, b.sum_S)
    print(N, r.kernel, r.count, "ok")
e.close()
