#!/usr/bin/env python
"""Time the fp64 kernel (whole runs and as the fix-up of the tensor path) for every library under devlibs/*/.
    python tools/exact_bench.py [N ...]"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))


def child(N):
    import numpy as np
    from mcr_b200 import api
    from mcr_b200.synthetic import synthetic_model
    m = synthetic_model(N)
    e = api.Engine(0)
    e.upload_sigma(m.mu, m.Sigma)
    Mx = max(100_000, int(2e11 / (N * N)) // 1000 * 1000)
    Ma = 20 * Mx
    out = {}
    for name, kern, M in (("exact", api.KERNEL_EXACT, Mx), ("auto", api.KERNEL_AUTO, Ma)):
        for _ in range(2):
            e.run(M, 67, m.x, m.k, kernel=kern)
        ts = []
        for _ in range(5):
            r = e.run(M, 67, m.x, m.k, kernel=kern)
            ts.append(r.device_ms)
        out[name] = {"M": M, "ms": float(np.median(ts)), "count": r.count, "sum_S": r.sum_S, "rechecked": r.n_rechecked}
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
        sys.exit(0)
    Ns = [int(a) for a in sys.argv[1:]] or [160, 600, 1000]
    libs = sorted((ROOT / "devlibs").glob("*/libmcr_b200.so"))
    for N in Ns:
        ref = None
        for lib in libs:
            env = dict(os.environ, MCR_B200_LIB=str(lib))
            p = subprocess.run([sys.executable, __file__, "--child", str(N)], env=env, capture_output=True, text=True)
            if p.returncode != 0:
                print(f"N={N} {lib.parent.name:10s} FAILED: {p.stderr.strip().splitlines()[-1] if p.stderr.strip() else p.returncode}")
                continue
            d = json.loads(p.stdout.strip().splitlines()[-1])
            key = (d["exact"]["count"], d["exact"]["sum_S"], d["auto"]["count"], d["auto"]["sum_S"])
            ref = ref or key
            x, a = d["exact"], d["auto"]
            print(f"N={N:5d} {lib.parent.name:10s} exact M={x['M']:8d} {x['ms']:9.2f} ms {x['M'] / x['ms'] * 1e3:10.3e} sims/s | "
                  f"auto M={a['M']:9d} {a['ms']:9.2f} ms {a['M'] / a['ms'] * 1e3:10.3e} sims/s rechecked {a['rechecked']:7d}  "
                  f"{'ok' if key == ref else 'MISMATCH'}", flush=True)
