#!/usr/bin/env python
"""Row-slab kernel vs single-pass wide kernel for N > 128 (env MCR_LARGE_N picks the path; one subprocess per path).
    python tools/large_n_sweep.py [N ...]"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))


def child(Ns):
    import numpy as np
    from mcr_b200 import api
    from mcr_b200.synthetic import synthetic_model
    e = api.Engine(0)
    for N in Ns:
        m = synthetic_model(N)
        e.upload_sigma(m.mu, m.Sigma)
        M = max(1_000_000, int(4e12 / (N * N)) // 100_000 * 100_000)
        for _ in range(2):
            e.run(M, 67, m.x, m.k)
        ts = []
        for _ in range(5):
            r = e.run(M, 67, m.x, m.k)
            ts.append(r.device_ms)
        print(json.dumps({"N": N, "M": M, "ms": float(np.median(ts)), "count": r.count, "sum_S": r.sum_S, "rechecked": r.n_rechecked,
                          "NT": r.padded_N, "launches": r.launches}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child([int(a) for a in sys.argv[2:]])
        sys.exit(0)
    Ns = [a for a in sys.argv[1:]] or ["160", "256", "320", "384", "512", "600", "640", "768", "896", "1000"]
    res = {}
    for path in ("slab", "wide"):
        env = dict(os.environ, MCR_LARGE_N=path)
        p = subprocess.run([sys.executable, __file__, "--child"] + Ns, env=env, capture_output=True, text=True)
        if p.returncode != 0:
            print(path, "FAILED", p.stderr[-800:])
            continue
        res[path] = {d["N"]: d for d in map(json.loads, p.stdout.strip().splitlines())}
    for N in map(int, Ns):
        a, b = res.get("slab", {}).get(N), res.get("wide", {}).get(N)
        if a and b:
            same = (a["count"], a["sum_S"]) == (b["count"], b["sum_S"])
            print(f"N={N:5d} M={a['M']:9d}  slab {a['ms']:8.2f} ms {a['M'] / a['ms'] * 1e3:10.3e} sims/s | wide {b['ms']:8.2f} ms "
                  f"{b['M'] / b['ms'] * 1e3:10.3e} sims/s  wide/slab speed {a['ms'] / b['ms']:5.2f}  counts {'equal' if same else 'DIFFER'}")
