#!/bin/bash
# Development build that recompiles ONE translation unit of csrc/ with extra flags and links it with the objects of the regular
# build (gpu-montecarlo-risk_b200/build/, `make lib` first) into devlibs/<name>/ for tools/tune.py A/B runs:
#   tools/devbuild_one.sh nomma mcr_stream -DMCR_STREAM_ABLATE=3
#   tools/devbuild_one.sh z4e2 mcr_wide -DMCR_WIDE_ELAG=2 -DMCR_WIDE_ZSLOTS_96=4 -DMCR_WIDE_RING_96=3
set -e
name=$1; unit=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
pkg=$root/gpu-montecarlo-risk_b200
out=$root/devlibs/$name
mkdir -p "$out"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c "$pkg/csrc/$unit.cu" -o "$out/$unit.o"
objs=$(ls "$pkg"/build/*.o | grep -v "/$unit.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libmcr_b200.so" $objs "$out/$unit.o" -ldl
rm -f "$out/$unit.o"
echo "built $out/libmcr_b200.so"
