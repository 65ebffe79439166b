#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`): headline metrics + stall samples by opcode.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [--window START COUNT]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_elapsed",
        "idc__request_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + (["--print-source", "sass"] if page == "source" else []),
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = ncu(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))
    for k in KEYS:
        if k in d:
            print(f"{k:75s} {d[k][1]:>16s} {d[k][0]}")
    for k, (u, v) in d.items():
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(v or 0) > 0.05:
            print(f"  {k.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ''):40s} {float(v):.3f}")
    rows = ncu(rep, "source")
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) >= len(hdr)]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    by_op = collections.defaultdict(collections.Counter)
    execd = collections.Counter()
    for r in data:
        src = r[ix["Source"]].strip().split()
        op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
        execd[op] += int(r[ix["Instructions Executed"]] or 0)
        by_op[op]["samples"] += int(r[ix["# Samples"]] or 0)
        for c in stall_cols:
            by_op[op][c] += int(r[ix[c]] or 0)
    S = sum(v["samples"] for v in by_op.values()) or 1
    E = sum(execd.values()) or 1
    print(f"{'op':8s} {'samples':>8s} {'exec':>7s}  top stall reasons (share of the op's samples)")
    for op, c in sorted(by_op.items(), key=lambda kv: -kv[1]["samples"])[:12]:
        s = c["samples"] or 1
        top = ", ".join(f"{k[6:]}={v / s:.2f}" for k, v in c.most_common(5) if k != "samples")
        print(f"{op:8s} {c['samples'] / S:8.3f} {execd[op] / E:7.3f}  {top}")
    if "--window" in sys.argv:
        a = int(sys.argv[sys.argv.index("--window") + 1]); n = int(sys.argv[sys.argv.index("--window") + 2])
        hot = [i for i, r in enumerate(data) if int(r[ix["Instructions Executed"]] or 0) > 1000]
        for r in data[hot[0] + a: hot[0] + a + n]:
            print(f"{r[ix['Source']].strip()[:76]:78s} smp={r[ix['# Samples']]:>5s} ssb={r[ix['stall_short_sb']]:>4s} wait={r[ix['stall_wait']]:>4s} "
                  f"math={r[ix['stall_math']]:>4s} disp={r[ix['stall_dispatch']]:>4s} nsel={r[ix['stall_not_selected']]:>4s}")


if __name__ == "__main__":
    main()
