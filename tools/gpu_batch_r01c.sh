#!/bin/bash
# One gpurun call: parity tests on the in-tree build, A/B timings (unroll variants; wide-kernel ablations), bench lines, launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest default rc=$?" | tee $O/summary.txt
MCR_TUNE_LIBS=base,cur,tc_u1,tc_u3 timeout 300 python tools/tune.py 50 100000000 > $O/tune_n50.log 2>&1; cat $O/tune_n50.log | tee -a $O/summary.txt
MCR_TUNE_LIBS=cur,tc_u1,tc_u3 timeout 300 python tools/tune.py 100 100000000 > $O/tune_n100.log 2>&1; cat $O/tune_n100.log | tee -a $O/summary.txt
for n in 600 320; do
  MCR_TUNE_NORECHECK=1 MCR_TUNE_LIBS=cur,wide_ablate1,wide_ablate2,wide_ablate3 timeout 300 python tools/tune.py $n 20000000 > $O/tune_wide_n$n.log 2>&1; echo "N=$n" | tee -a $O/summary.txt; cat $O/tune_wide_n$n.log | tee -a $O/summary.txt
done
MCR_LARGE_N=wide MCR_TUNE_NORECHECK=1 MCR_TUNE_LIBS=cur,wide_ablate1,wide_ablate2,wide_ablate3 timeout 300 python tools/tune.py 1000 10000000 > $O/tune_wide_n1000.log 2>&1; echo "N=1000" | tee -a $O/summary.txt; cat $O/tune_wide_n1000.log | tee -a $O/summary.txt
timeout 400 python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 400 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench ref rc=$?" | tee -a $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
tail -n 3 $O/pytest_gpu.log
