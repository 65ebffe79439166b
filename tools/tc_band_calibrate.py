#!/usr/bin/env python
"""How much margin does the tensor-core guard band have?  Runs the tensor kernel with the band scaled
down (MCR_TC_BAND_SCALE, read once per process -> one child per scale) and compares count / sum_S with
the fp64 exact kernel on the same Philox stream.  The band is safe while the results stay identical."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))

def child(N, M, kernel):
    from mcr_b200 import api
    from mcr_b200.synthetic import synthetic_model
    m = synthetic_model(N); e = api.Engine(0); e.upload_sigma(m.mu, m.Sigma)
    r = e.run(M, 4242, m.x, m.k, kernel={"exact": api.KERNEL_EXACT, "tensor": api.KERNEL_TENSOR, "ffma": api.KERNEL_FFMA}[kernel])
    print(json.dumps({"count": r.count, "sum_S": r.sum_S, "rechecked": r.n_rechecked, "ms": r.device_ms}))

if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]); sys.exit(0)
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    Ns = [int(a) for a in sys.argv[2:]] or [50, 100, 128, 25]
    for N in Ns:
        def run(kernel, scale=None):
            env = dict(os.environ)
            if scale is not None: env["MCR_TC_BAND_SCALE"] = str(scale)
            p = subprocess.run([sys.executable, __file__, "--child", str(N), str(M), kernel], env=env, capture_output=True, text=True)
            return json.loads(p.stdout.strip().splitlines()[-1])
        ex = run("exact")
        print(f"N={N} M={M} exact: count={ex['count']} sum_S={ex['sum_S']} ({ex['ms']:.0f} ms)")
        for scale in (1, 0.5, 0.25, 0.125, 0.0625, 0.03125, 0.015625, 0.0078125, 1e-9):
            r = run("tensor", scale)
            ok = (r["count"], r["sum_S"]) == (ex["count"], ex["sum_S"])
            print(f"   band x{scale:<10g} rechecked={r['rechecked']:>8d}  d_count={r['count'] - ex['count']:+d} d_sumS={r['sum_S'] - ex['sum_S']:+d}  {'exact' if ok else 'DIFFERS'}")
