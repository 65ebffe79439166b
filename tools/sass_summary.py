#!/usr/bin/env python
"""Per-kernel SASS opcode summary of the built library (cuobjdump -sass), for profiles/: how many tcgen05 MMA (UTCHMMA),
TMEM load/store (LDTM / STTM), bulk-copy (UBLKCP), packed-FP32 (FFMA2), MUFU, IMAD.WIDE instructions each kernel holds, and
that none of them touches local memory (LDL / STL = spills).
    python tools/sass_summary.py [libmcr_b200.so] > profiles/rNN_sass_opcode_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpu-montecarlo-risk_b200" / "libmcr_b200.so"
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FFMA", "DFMA", "DMUL", "DADD", "IMAD.WIDE", "LOP3", "MUFU",
       "REDUX", "LDC", "LDCU", "LDS", "STS", "LDG", "STG", "ATOM", "ATOMS", "RED", "LDL", "STL", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
demangle = lambda names: subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                kernels[cur][o] += 1
names = demangle(list(kernels))
print(f"# SASS opcode counts per kernel of {lib.name} (static instruction counts; cuobjdump -sass, sm_100a)")
print("# columns: total | " + " ".join(OPS))
agg = collections.OrderedDict()
for (mangled, c), name in zip(kernels.items(), names):
    short = re.sub(r"\((?:unsigned )?(?:int|bool|long)\)", "", name)          # template arguments print as (int)64
    short = re.sub(r"\(.*", "", short).replace("void ", "").replace("mcr::", "")
    agg[short] = c
for short, c in sorted(agg.items()):
    cols = " ".join(f"{o}={c[o]}" for o in OPS if c[o])
    print(f"{short:70s} total={c['_total']:6d}  {cols}")
spill = [k for k, c in agg.items() if c["LDL"] or c["STL"]]
print(f"# kernels with local-memory instructions (LDL/STL): {spill if spill else 'none'}")
print(f"# kernels issuing tcgen05.mma (UTCHMMA): {sorted(k for k, c in agg.items() if c['UTCHMMA'])}")
