#!/bin/bash
# after the MMA-issue fix: does the Z depth / early epilogue matter now (b0 = 2 Z slots, whole epilogue at the boundary; e2 = ELAG 2;
# cur = committed default ELAG 3), and where do the single-pass and the row-slab kernel cross (N = 256, 768)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
export MCR_TUNE_CHILD_TIMEOUT=40
: > $O/summary_j.txt
for n in 320 600; do
  echo "## N=$n M=20000000" >> $O/summary_j.txt
  MCR_TUNE_LIBS=cur,b0,e2 timeout 90 python tools/tune.py $n 20000000 >> $O/summary_j.txt 2>&1
done
for nm in "256 20000000" "768 8000000"; do
  set -- $nm
  for k in wide slab; do
    echo "## N=$1 M=$2 MCR_LARGE_N=$k" >> $O/summary_j.txt
    MCR_LARGE_N=$k MCR_TUNE_LIBS=cur timeout 60 python tools/tune.py $1 $2 >> $O/summary_j.txt 2>&1
  done
done
for v in b0 e2; do
  MCR_B200_LIB=$PWD/devlibs/$v/libmcr_b200.so timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_n or several_tiles or slab_fixup or histogram" > $O/pytest_$v.log 2>&1; echo "pytest $v rc=$? $(tail -n 1 $O/pytest_$v.log)" >> $O/summary_j.txt
done
cat $O/summary_j.txt
