#!/bin/bash
# four GPUs: weak scaling of the batch decode (1024 streams per GPU) with the device-record frame gather
mkdir -p gpurun_out
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu --no-e2e) > gpurun_out/c10_bench_4gpu.log 2>&1
grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"phases_ms": {[^}]*}\|"rank_spread": {[^}]*}' gpurun_out/c10_bench_4gpu.log | head; tail -n 3 gpurun_out/c10_bench_4gpu.log | cut -c1-300
