#!/bin/bash
# One parametrised GPU batch (replaces the per-experiment gpu_batch_r01*.sh of round 1).  Run on a B200 box with
#   gpurun --timeout 900 -- 'bash tools/gpu_batch.sh <tag> <stage> [<stage> ...]'
# Every stage writes into gpurun_out/ (merged back by gpurun) and appends "<stage> rc=N" to gpurun_out/summary_<tag>.txt.
# Stages:
#   tests[:EXPR]          pytest -m gpu (optionally -k EXPR)
#   testsq:EXPR           the same with -k EXPR under a short leash (new kernels: 120 s per test, 300 s in all)
#   smoke                 __graft_entry__.smoke()
#   bench[:ARGS]          python bench.py ARGS            (1 GPU; ARGS with ',' for spaces)
#   benchN:<n>[:ARGS]     torchrun --nproc-per-node n bench.py --gpus n ARGS
#   refbench              python bench.py --impl reference --steps 3 --warmup 1
#   launches              ncu launch list of the default bench command (gpu__time_duration.sum, no clock control)
#   ncu:<regex>:<N>:<M>   ncu --set full capture of one launch of the kernel matching <regex> at N assets, M trials
#   tune:<N>:<M>[:libs]   tools/tune.py over devlibs/*/ (or the comma-separated subset)
#   tuneraw:<N>:<M>[:libs] the same with no_recheck (ablated builds return garbage decisions)
#   ubench:<name>         tools/<name> (a prebuilt micro-benchmark binary)
#   env:VAR=VALUE         export a variable for the stages that follow
#   py:<script>[:ARGS]    python tools/<script> ARGS
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
tag=$1; shift
S=$O/summary_$tag.txt; : > $S
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > $O/gpu_$tag.txt 2>&1
for stage in "$@"; do
  IFS=':' read -r kind a b c <<< "$stage"
  case $kind in
    tests)    ( time timeout 900 python -m pytest tests -m gpu -x -q ${a:+-k "$a"} ) > $O/pytest_$tag.log 2>&1; rc=$?; tail -n 6 $O/pytest_$tag.log ;;
    testsq)   ( time timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 -k "$a" ) > $O/pytestq_$tag.log 2>&1; rc=$?; tail -n 15 $O/pytestq_$tag.log ;;
    smoke)    timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$tag.log 2>&1; rc=$?; tail -n 3 $O/smoke_$tag.log ;;
    bench)    timeout 600 python bench.py ${a//,/ } > $O/bench_$tag.json 2> $O/bench_$tag.err; rc=$?; cut -c1-400 $O/bench_$tag.json; tail -n 3 $O/bench_$tag.err ;;
    benchN)   timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $a --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $a ${b//,/ } \
                > $O/bench_${tag}_${a}gpu.json 2> $O/bench_${tag}_${a}gpu.err; rc=$?; tail -n 1 $O/bench_${tag}_${a}gpu.json | cut -c1-400; tail -n 3 $O/bench_${tag}_${a}gpu.err ;;
    refbench) timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$tag.json 2> $O/bench_ref_$tag.err; rc=$?; cut -c1-300 $O/bench_ref_$tag.json ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$tag.csv \
                python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu_$tag.log 2>&1; rc=$? ;;
    ncu)      timeout 600 ncu --set full --clock-control none --import-source on -k regex:$a -s 1 -c 1 -f -o $O/${tag}_${a}_n$b \
                python tools/prof_one.py $b auto $c 2 > $O/ncu_${tag}_${a}_n$b.log 2>&1; rc=$?; tail -n 2 $O/ncu_${tag}_${a}_n$b.log ;;
    tune)     echo "## N=$a M=$b" >> $S; MCR_TUNE_LIBS=$c timeout 600 python tools/tune.py $a $b >> $S 2>&1; rc=$? ;;
    tuneraw)  echo "## (no recheck) N=$a M=$b" >> $S; MCR_TUNE_NORECHECK=1 MCR_TUNE_LIBS=$c timeout 600 python tools/tune.py $a $b >> $S 2>&1; rc=$? ;;
    ubench)   timeout 300 tools/$a > $O/${a}_$tag.log 2>&1; rc=$?; tail -n 40 $O/${a}_$tag.log ;;
    py)       timeout 900 python tools/$a ${b//,/ } > $O/${a%.py}_$tag.log 2>&1; rc=$?; tail -n 30 $O/${a%.py}_$tag.log ;;
    env)      export "$a"; rc=0 ;;
    *)        echo "unknown stage $stage"; rc=99 ;;
  esac
  echo "$stage rc=$rc" | tee -a $S
done
cat $S
