#!/bin/bash
# single-pass kernel: Z-slot depth / transform-ring depth / early slab epilogue variants (tools/devbuild_wide.sh) against prev
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
export MCR_TUNE_CHILD_TIMEOUT=25
: > $O/summary_g.txt
for n in 320 600; do
  echo "## N=$n" >> $O/summary_g.txt
  MCR_TUNE_LIBS=prev,v0,vA,vB,vC,vD,vE timeout 150 python tools/tune.py $n 20000000 >> $O/summary_g.txt 2>&1
done
echo "## N=1000 MCR_LARGE_N=wide (prev,vA,vD) then slab (prev)" >> $O/summary_g.txt
MCR_LARGE_N=wide MCR_TUNE_LIBS=prev,vA,vD timeout 80 python tools/tune.py 1000 8000000 >> $O/summary_g.txt 2>&1
MCR_LARGE_N=slab MCR_TUNE_LIBS=prev timeout 40 python tools/tune.py 1000 8000000 >> $O/summary_g.txt 2>&1
( time MCR_B200_LIB=$PWD/devlibs/vB/libmcr_b200.so timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_n or several_tiles or slab_fixup or histogram" ) > $O/pytest_vB.log 2>&1; echo "pytest vB rc=$?" >> $O/summary_g.txt
cat $O/summary_g.txt; tail -n 5 $O/pytest_vB.log
