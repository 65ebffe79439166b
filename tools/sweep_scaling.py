#!/usr/bin/env python
"""BASELINE config 5: asset count x path count sweep, paths sharded over all ranks (strong scaling of a fixed
total M), one 64-byte all-reduce per run.  Launch: torchrun --nproc-per-node G tools/sweep_scaling.py
Also prints BASELINE config 3 (N=100, M=1e9).  Device time = CUDA events on the launching stream, max over ranks."""
import json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))
import torch, torch.distributed as dist
from mcr_b200 import api, model_io
from mcr_b200.synthetic import synthetic_model

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = api.Engine(local)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
partials = torch.zeros(api.MCR_PARTIAL_WORDS, dtype=torch.int64, device="cuda")
# rough single-GPU rates (sims/s) to keep each point under a few seconds
RATE = {16: 5e10, 32: 2e10, 64: 1.2e10, 100: 8e9, 128: 5e9, 256: 2e9, 512: 6e8, 600: 5e8, 1024: 2.4e8}
rows = []
for N in (16, 32, 64, 128, 256, 512, 1024, 100, 600):
    m = synthetic_model(N)
    eng.upload_sigma(m.mu, m.Sigma)
    for M in ((10**9,) if N == 100 else (10**8,) if N == 600 else (10**7, 10**8, 10**9, 10**10)):
        if M / (RATE[N] * world) > 4.0:
            continue
        off, cnt = model_io.shard_ranges(M, world)[rank]
        def run():
            eng.run_async(cnt, 67, m.x, m.k, partials.data_ptr(), stream.cuda_stream, path_offset=off)
            if world > 1:
                dist.all_reduce(partials)
        run(); torch.cuda.synchronize()
        if world > 1: dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps): run()
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res = api.result_from_partials(partials.cpu().numpy().astype(np.uint64))
        assert res.n_paths == M
        if rank == 0:
            row = {"N": N, "M": M, "gpus": world, "ms": float(ms.item()), "sims_per_s": M / float(ms.item()) * 1e3,
                   "tflops_equiv": M * (N * N + 3 * N) / float(ms.item()) * 1e3 / 1e12, "count": res.count, "P_hat": res.P_hat,
                   "rechecked": res.n_rechecked}
            rows.append(row); print(json.dumps(row), flush=True)
if world > 1:
    dist.destroy_process_group()
