#!/usr/bin/env python
"""Timeline of one tile of the streamed kernel from a -DMCR_STREAM_TRACE=1 build (tools/devbuild_one.sh trace mcr_stream -DMCR_STREAM_TRACE=1):
    MCR_B200_LIB=devlibs/trace/libmcr_b200.so python tools/stream_trace.py [N] [tile] [supplied]"""
import ctypes as C
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))
from mcr_b200 import api, _lib
from mcr_b200.synthetic import synthetic_model

N = int(sys.argv[1]) if len(sys.argv) > 1 else 600
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 6
supplied = len(sys.argv) > 3 and sys.argv[3] == "supplied"   # caller-supplied normals (TMA boxes) instead of Philox
m = synthetic_model(N)
e = api.Engine(0)
e.upload_sigma(m.mu, m.Sigma)
M = 148 * 128 * 12
if supplied:
    import torch
    Zd = torch.randn(M, N, dtype=torch.float64, device="cuda")
    run = lambda: e.run_with_normals_device(Zd.data_ptr(), M, m.x, m.k, kernel=api.KERNEL_TENSOR, no_recheck=True)
else:
    run = lambda: e.run(M, 67, m.x, m.k, no_recheck=True)
run()
lib = _lib.load()
ROLES, CAP = 4, 16384
buf = np.zeros(ROLES * CAP, np.uint64)
cnt = np.zeros(ROLES, np.uint32)
lib.mcr_stream_trace_dump(buf.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 1)   # drop the warm-up run
run()
assert lib.mcr_stream_trace_dump(buf.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 1) == 0
names = {0: "gen w0 ", 1: "gen w12", 2: "mma    ", 3: "loader "}
tags = {0: {1: "step begins", 2: "normals ready", 3: "slot free", 4: "handed over", 5: "epilogue begins", 6: "epilogue done", 7: "pass MMAs complete"},
        1: {1: "step begins", 2: "normals ready", 3: "slot free", 4: "handed over"},
        2: {1: "wait slot", 2: "slot ready", 3: "accumulators free", 4: "piece arrived", 5: "piece issued"},
        3: {1: "copy issued"}}
ev = []
for r in range(ROLES):
    for v in buf[r * CAP: r * CAP + int(cnt[r])]:
        v = int(v)
        kb = (v >> 40) & 0xFFFF
        kb = kb - 65536 if kb >= 32768 else kb
        ev.append((v & 0xFFFFFFFFFF, r, v >> 56, kb))
ev.sort()
# tiles of CTA 0 are delimited by the generator's "step begins" with kb = -1 in pass 0: every (npass)-th such event
starts = [t for t, r, tag, kb in ev if r == 0 and tag == 1 and kb == -1]
npass = 1 if N <= 384 else (2 if N <= 768 else 3)
t0, t1 = starts[tile * npass], starts[(tile + 1) * npass]
print(f"N={N}: tile {tile} of CTA 0 spans {t1 - t0} clocks ({cnt} events recorded)")
last = {}
for t, r, tag, kb in ev:
    if t0 <= t < t1 + 2000:
        dt = t - last.get(r, t)
        last[r] = t
        print(f"{t - t0:8d}  {names[r]}  kb={kb:3d}  {tags[r].get(tag, tag):20s} (+{dt})")
