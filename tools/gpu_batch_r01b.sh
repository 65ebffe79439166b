#!/bin/bash
# One gpurun call: parity tests on the in-tree build, A/B timing of the development builds, bench lines, ncu captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest default rc=$?" | tee -a $O/summary.txt
timeout 300 python tools/tune.py 50 100000000 > $O/tune_n50.log 2>&1; cat $O/tune_n50.log | tee -a $O/summary.txt
timeout 300 python tools/tune.py 100 100000000 > $O/tune_n100.log 2>&1; cat $O/tune_n100.log | tee -a $O/summary.txt
( time MCR_B200_LIB=$PWD/devlibs/fold/libmcr_b200.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "philox_run or every_tail or ragged or portfolio" ) > $O/pytest_fold.log 2>&1; echo "pytest fold rc=$?" | tee -a $O/summary.txt
for v in new fold; do
  MCR_B200_LIB=$PWD/devlibs/$v/libmcr_b200.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_path_kernel -s 1 -c 1 -f -o $O/r01_tc_n50_$v \
      python tools/prof_one.py 50 tensor 20000000 2 > $O/ncu_$v.log 2>&1; echo "ncu $v rc=$?" | tee -a $O/summary.txt
done
timeout 300 python bench.py --no-cpu-baseline > $O/bench_new.json 2> $O/bench_new.err; echo "bench new rc=$?" | tee -a $O/summary.txt
MCR_B200_LIB=$PWD/devlibs/fold/libmcr_b200.so timeout 300 python bench.py --no-cpu-baseline > $O/bench_fold.json 2> $O/bench_fold.err; echo "bench fold rc=$?" | tee -a $O/summary.txt
timeout 300 python tools/tune.py 64 50000000 > $O/tune_n64.log 2>&1; cat $O/tune_n64.log | tee -a $O/summary.txt
timeout 300 python tools/tune.py 128 50000000 > $O/tune_n128.log 2>&1; cat $O/tune_n128.log | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.log $O/pytest_fold.log
