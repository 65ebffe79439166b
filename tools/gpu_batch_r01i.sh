#!/bin/bash
# MMA issue from uniform registers (shuffled warp index, elect.sync leader, descriptors by addition) in the single-pass, row-slab
# and path kernels: timings with tools/tune.py (compare with profiles/r01_wide_zslots_early_epilogue.log), then the closing checks
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
export MCR_TUNE_CHILD_TIMEOUT=40 MCR_TUNE_LIBS=u2
: > $O/summary_i.txt
for nm in "50 100000000" "100 100000000" "160 20000000" "320 20000000" "600 20000000" "1000 8000000"; do
  set -- $nm; echo "## N=$1 M=$2" >> $O/summary_i.txt
  timeout 60 python tools/tune.py $1 $2 >> $O/summary_i.txt 2>&1
done
echo "## N=1000 M=8000000 MCR_LARGE_N=wide" >> $O/summary_i.txt
MCR_LARGE_N=wide timeout 60 python tools/tune.py 1000 8000000 >> $O/summary_i.txt 2>&1
( time timeout 240 python -m pytest tests -m gpu -x -q ) > $O/pytest_final.log 2>&1; echo "pytest rc=$?" >> $O/summary_i.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1; echo "smoke rc=$?" >> $O/summary_i.txt
timeout 200 python bench.py > $O/bench_r01i.json 2> $O/bench_r01i.err; echo "bench rc=$?" >> $O/summary_i.txt
timeout 120 python bench.py --assets 600 --paths 20000000 --steps 5 --no-cpu-baseline > $O/bench_r01i_n600.json 2> $O/bench_r01i_n600.err; echo "bench n600 rc=$?" >> $O/summary_i.txt
cat $O/summary_i.txt; tail -n 5 $O/pytest_final.log; tail -n 3 $O/smoke_final.log; cut -c1-200 $O/bench_r01i.json; cut -c1-200 $O/bench_r01i_n600.json
