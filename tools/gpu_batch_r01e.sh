#!/bin/bash
# 8-GPU call: bench.py at 8 and 4 ranks, strong-scaling sweep at 8 ranks
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench8 rc=$?" | tee $O/summary_8gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/sweep_scaling.py > $O/sweep_8gpu.jsonl 2> $O/sweep_8gpu.err; echo "sweep8 rc=$?" | tee -a $O/summary_8gpu.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 > $O/bench_4gpu.json 2> $O/bench_4gpu.err; echo "bench4 rc=$?" | tee -a $O/summary_8gpu.txt
tail -n 1 $O/bench_8gpu.json | cut -c1-200; tail -n 1 $O/bench_4gpu.json | cut -c1-200
