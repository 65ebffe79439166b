#!/bin/bash
# Development build of libmcr_b200.so with extra nvcc flags into devlibs/<name>/ (for tools/tune.py A/B runs):
#   tools/devbuild.sh fold -DMCR_TC_FOLD=1
# devlibs/ is git-ignored but travels with gpurun.  Only the tensor-kernel instantiations get the extra flags'
# worth of recompilation; everything is compiled in parallel.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
pkg=$root/gpu-montecarlo-risk_b200
out=$root/devlibs/$name
obj=$out/obj
mkdir -p "$obj"
flags="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $*"
pids=()
for f in mcr_api mcr_host mcr_exact mcr_tc mcr_wide mcr_stream mcr_stats; do
  nvcc $flags -c "$pkg/csrc/$f.cu" -o "$obj/$f.o" & pids+=($!)
done
for g in 0 1 2 3 4 5 6 7; do
  nvcc $flags -DMCR_INST_GROUP=$g -c "$pkg/csrc/mcr_ffma_inst.cu" -o "$obj/mcr_ffma_inst_$g.o" & pids+=($!)
done
for v in 0 1 2 3 4 5 6 7 8 9 10; do
  nvcc $flags -DMCR_TC_INST=$v -c "$pkg/csrc/mcr_tc_inst.cu" -o "$obj/mcr_tc_inst_$v.o" & pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libmcr_b200.so" "$obj"/*.o -ldl
rm -rf "$obj"
echo "built $out/libmcr_b200.so"
