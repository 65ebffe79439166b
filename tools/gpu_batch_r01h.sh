#!/bin/bash
# round-1 closing check of the committed default build: GPU tests, smoke, bench line, ncu capture of the single-pass kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
( time timeout 240 python -m pytest tests -m gpu -x -q ) > $O/pytest_final.log 2>&1; echo "pytest rc=$?" | tee $O/summary_h.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary_h.txt
timeout 200 python bench.py > $O/bench_r01h.json 2> $O/bench_r01h.err; echo "bench rc=$?" | tee -a $O/summary_h.txt
timeout 120 python bench.py --assets 600 --paths 20000000 --steps 5 --no-cpu-baseline > $O/bench_r01h_n600.json 2> $O/bench_r01h_n600.err; echo "bench n600 rc=$?" | tee -a $O/summary_h.txt
timeout 150 ncu --set full --clock-control none --import-source on -k regex:tc_wide -s 1 -c 1 -f -o $O/r01_wide_n600 python tools/prof_one.py 600 auto 2000000 2 > $O/ncu_wide.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary_h.txt
tail -n 3 $O/pytest_final.log; tail -n 3 $O/smoke_final.log; cut -c1-300 $O/bench_r01h.json; cut -c1-300 $O/bench_r01h_n600.json; tail -n 3 $O/ncu_wide.log
