#!/bin/bash
# Development build that recompiles only csrc/mcr_wide.cu with extra flags and links it with the objects of the regular
# build (gpu-montecarlo-risk_b200/build/, `make lib` first) into devlibs/<name>/ for tools/tune.py A/B runs:
#   tools/devbuild_wide.sh z4e2 -DMCR_WIDE_ELAG=2 -DMCR_WIDE_ZSLOTS_96=4 -DMCR_WIDE_RING_96=3
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
pkg=$root/gpu-montecarlo-risk_b200
out=$root/devlibs/$name
mkdir -p "$out"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c "$pkg/csrc/mcr_wide.cu" -o "$out/mcr_wide.o"
objs=$(ls "$pkg"/build/*.o | grep -v mcr_wide.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libmcr_b200.so" $objs "$out/mcr_wide.o" -ldl
rm -f "$out/mcr_wide.o"
echo "built $out/libmcr_b200.so"
