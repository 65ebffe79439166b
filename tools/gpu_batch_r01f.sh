#!/bin/bash
# single-pass kernel with per-slab epilogue inside the K loop (devlibs/w2) against the previous one (devlibs/prev)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
( time MCR_B200_LIB=$PWD/devlibs/w2/libmcr_b200.so timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_n or several_tiles or slab_fixup or histogram" ) > $O/pytest_w2.log 2>&1; echo "pytest w2 rc=$?" | tee $O/summary_w2.txt
for n in 320 512 600; do
  MCR_TUNE_LIBS=prev,w2 timeout 60 python tools/tune.py $n 20000000 > $O/tune_w2_n$n.log 2>&1; echo "N=$n" | tee -a $O/summary_w2.txt; cat $O/tune_w2_n$n.log | tee -a $O/summary_w2.txt
done
tail -n 4 $O/pytest_w2.log
