#!/usr/bin/env python
"""Single-process multi-GPU path (mcr_simulate_host with n_gpus > 1: NCCL all-reduce of the partial
words via dlopen'd libnccl) against the single-GPU result."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model
n = api.device_count()
print("devices:", n)
for N, M in ((50, 100_000_001), (100, 50_000_000)):
    m = synthetic_model(N); A = api.build_transform(m.Sigma, 0)
    one = api.simulate_host(m.mu, A, M, 67, m.x, m.k, a_is_lower=True, n_gpus=1)
    for g in [g for g in (2, 4, 8) if g <= n]:
        api.simulate_host(m.mu, A, M, 67, m.x, m.k, a_is_lower=True, n_gpus=g)  # warm (NCCL init)
        t0 = time.perf_counter()
        r = api.simulate_host(m.mu, A, M, 67, m.x, m.k, a_is_lower=True, n_gpus=g)
        dt = time.perf_counter() - t0
        ok = (r.count, r.sum_S, r.n_paths) == (one.count, one.sum_S, one.n_paths)
        print(f"N={N} M={M} n_gpus={g}: count={r.count} sum_S={r.sum_S} device_ms={r.device_ms:.3f} wall_ms={dt*1e3:.3f} "
              f"(1 GPU device_ms={one.device_ms:.3f}) {'IDENTICAL' if ok else 'MISMATCH'}")
    # extension outputs (histogram of S, portfolio sums / histogram): one more all-reduce per buffer
    e1 = api.simulate_host(m.mu, A, 5_000_001, 67, m.x, m.k, a_is_lower=True, n_gpus=1, portfolio_threshold=0.015, want_hist=True)
    for g in [g for g in (2, 4, 8) if g <= n]:
        eg = api.simulate_host(m.mu, A, 5_000_001, 67, m.x, m.k, a_is_lower=True, n_gpus=g, portfolio_threshold=0.015, want_hist=True)
        same = (bool((eg.hist_S == e1.hist_S).all()) and eg.portfolio["hits"] == e1.portfolio["hits"]
                and bool((eg.portfolio["hist"] == e1.portfolio["hist"]).all()))
        print(f"   extension outputs n_gpus={g}: hist_S / portfolio hits / portfolio histogram {'IDENTICAL' if same else 'MISMATCH'}; "
              f"VaR99 {eg.portfolio['VaR'][0.99]:.6f} vs {e1.portfolio['VaR'][0.99]:.6f}, ES99 {eg.portfolio['ES'][0.99]:.6f} vs {e1.portfolio['ES'][0.99]:.6f}, "
              f"mean rel diff {abs(eg.portfolio['mean'] / e1.portfolio['mean'] - 1):.2e}")
    S = api.simulate_host(m.mu, A, 1_000_003, 67, m.x, m.k, a_is_lower=True, n_gpus=min(n, 2), want_S_values=True)
    S1 = api.simulate_host(m.mu, A, 1_000_003, 67, m.x, m.k, a_is_lower=True, n_gpus=1, want_S_values=True)
    print("   S_values across shards identical:", bool((S.S_values == S1.S_values).all()))
