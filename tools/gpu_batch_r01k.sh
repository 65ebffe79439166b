#!/bin/bash
# closing check of the default build (2 Z slots, epilogue at the tile boundary, MMA issue from uniform registers)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
( time timeout 100 python -m pytest tests -m gpu -x -q ) > $O/pytest_final.log 2>&1; echo "pytest rc=$?" | tee $O/summary_k.txt
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary_k.txt
timeout 60 python bench.py --assets 600 --paths 20000000 --steps 5 --no-cpu-baseline > $O/bench_r01k_n600.json 2> $O/bench_r01k_n600.err; echo "bench n600 rc=$?" | tee -a $O/summary_k.txt
tail -n 4 $O/pytest_final.log; tail -n 2 $O/smoke_final.log; cut -c1-160 $O/bench_r01k_n600.json
