#!/bin/bash
# 2-GPU call: single-process NCCL path (incl. extension outputs), torchrun bench and sweep at 2 ranks
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python tools/multi_gpu_check.py > $O/multi_gpu_check_2gpu.log 2>&1; echo "multi_gpu_check rc=$?" | tee $O/summary_2gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?" | tee -a $O/summary_2gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/sweep_scaling.py > $O/sweep_2gpu.jsonl 2> $O/sweep_2gpu.err; echo "sweep2 rc=$?" | tee -a $O/summary_2gpu.txt
cat $O/multi_gpu_check_2gpu.log; tail -n 2 $O/bench_2gpu.json | cut -c1-300
