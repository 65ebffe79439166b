#!/bin/bash
# Round-2 opener (prepared at the end of round 1, not yet run): (1) the tcgen05.mma retire-rate micro-benchmark,
# (2) the single-pass kernel with more than two Z slots WITHOUT the early epilogue, which the closing ncu capture suggests
# (generators wait 31 % of the samples for a free Z slot).  Build the variants first:
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gpu-montecarlo-risk_b200/csrc -o tools/ubench_umma tools/ubench_umma.cu
#   cp gpu-montecarlo-risk_b200/libmcr_b200.so devlibs/cur/   (after mkdir -p devlibs/cur)
#   tools/devbuild_wide.sh z3 -DMCR_WIDE_ZSLOTS_128=3 -DMCR_WIDE_RING_128=3 -DMCR_WIDE_ZSLOTS_96=3 -DMCR_WIDE_RING_96=4 -DMCR_WIDE_ZSLOTS_64=4 -DMCR_WIDE_RING_64=4
#   tools/devbuild_wide.sh z5 -DMCR_WIDE_ZSLOTS_128=3 -DMCR_WIDE_RING_128=3 -DMCR_WIDE_ZSLOTS_96=5 -DMCR_WIDE_RING_96=3 -DMCR_WIDE_ZSLOTS_64=6 -DMCR_WIDE_RING_64=3
#   tools/devbuild_wide.sh g768 -DMCR_WIDE_GEN_THREADS_96=768          (N = 513..640: all 24 generator warps busy, 8 parts of 8 columns)
#   tools/devbuild_wide.sh g768z4 -DMCR_WIDE_GEN_THREADS_96=768 -DMCR_WIDE_ZSLOTS_96=4 -DMCR_WIDE_RING_96=3
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
timeout 120 tools/ubench_umma > $O/ubench_umma.log 2>&1; echo "ubench rc=$?" | tee $O/summary_r02a.txt
export MCR_TUNE_CHILD_TIMEOUT=40
for n in 320 600; do
  echo "## N=$n M=20000000" >> $O/summary_r02a.txt
  MCR_TUNE_LIBS=cur,z3,z5,g768,g768z4 timeout 150 python tools/tune.py $n 20000000 >> $O/summary_r02a.txt 2>&1
done
cat $O/ubench_umma.log $O/summary_r02a.txt
