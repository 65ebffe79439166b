#!/usr/bin/env python
"""Bring-up of the row-slab tensor kernel (N > 128): exactness vs oracle on dumped normals + timing vs the fp64 kernel."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200")); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model
o = oracle_lib.Oracle(); e = api.Engine(0)
for N in [int(a) for a in sys.argv[1:]] or [160, 600]:
    m = synthetic_model(N); A = e.upload_sigma(m.mu, m.Sigma)
    M = 4000
    Z, _ = e.dump_normals(M, 67)
    ro = o.run_normals(A, m.mu, m.k, m.x, Z)
    for name, kw in (("tensor", {}), ("tensor-raw", dict(no_recheck=True)), ("exact", dict(kernel=api.KERNEL_EXACT))):
        r = e.run(M, 67, m.x, m.k, **kw)
        print(f"N={N} {name:10s} count={r.count} sumS={r.sum_S} rechecked={r.n_rechecked} kernel={r.kernel} NT={r.padded_N} launches={r.launches} "
              f"oracle=({ro['count']},{ro['sum_S']}) {'OK' if (r.count, r.sum_S) == (ro['count'], ro['sum_S']) else 'DIFF'}", flush=True)
    Mb = 20_000_000 if N <= 256 else 10_000_000
    for name, kern, Mk in (("tensor", api.KERNEL_AUTO, Mb), ("exact", api.KERNEL_EXACT, Mb // 20)):
        e.run(Mk, 67, m.x, m.k, kernel=kern)
        r = e.run(Mk, 67, m.x, m.k, kernel=kern)
        F = N * N + 3 * N
        print(f"   {name:7s} M={Mk}: run {r.device_ms:.2f} ms, {Mk / r.device_ms * 1e3:.3e} sims/s, {Mk * F / r.device_ms * 1e3 / 1e12:.1f} TFLOP/s-equiv, rechecked {r.n_rechecked}", flush=True)
