#!/usr/bin/env python
"""Bring-up check of the tensor-core kernel: exactness vs the oracle on dumped normals + timing."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200")); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model

o = oracle_lib.Oracle()
e = api.Engine(0)
Ns = [int(a) for a in sys.argv[1:]] or [50]
for N in Ns:
    m = synthetic_model(N)
    A = e.upload_sigma(m.mu, m.Sigma)
    M = 200000
    Z, _ = e.dump_normals(M, 67)
    ro = o.run_normals(A, m.mu, m.k, m.x, Z)
    for name, kw in (("tensor", dict(kernel=api.KERNEL_TENSOR)), ("tensor-raw", dict(kernel=api.KERNEL_TENSOR, no_recheck=True)),
                     ("ffma", dict(kernel=api.KERNEL_FFMA))):
        r = e.run(M, 67, m.x, m.k, **kw)
        print(f"N={N} {name:10s} count={r.count} sumS={r.sum_S} rechecked={r.n_rechecked} kernel={r.kernel} NT={r.padded_N} "
              f"oracle=({ro['count']},{ro['sum_S']}) {'OK' if (r.count, r.sum_S) == (ro['count'], ro['sum_S']) else 'DIFF'}", flush=True)
    h = api.simulate_host(m.mu, A, M, 67, m.x, m.k, a_is_lower=True, want_S_values=True, kernel=api.KERNEL_TENSOR)
    print("   S_values equal:", bool(np.array_equal(h.S_values, ro["S_values"])), "mismatches:", int((h.S_values != ro["S_values"]).sum()), flush=True)
    for name, kern in (("tensor", api.KERNEL_TENSOR), ("ffma", api.KERNEL_FFMA)):
        Mb = 50_000_000
        for _ in range(2):
            e.run(Mb, 67, m.x, m.k, kernel=kern)
        ks = []
        for _ in range(5):
            r = e.run(Mb, 67, m.x, m.k, kernel=kern)
            ks.append(e.last_main_kernel_ms()[0])
        kms = float(np.median(ks)); F = N * N + 3 * N
        print(f"   {name:7s} M={Mb}: kernel {kms:.3f} ms, run {r.device_ms:.3f} ms, {Mb / kms * 1e3:.3e} sims/s, {Mb * F / kms * 1e3 / 1e12:.1f} TFLOP/s-equiv, rechecked {r.n_rechecked}", flush=True)
