#!/bin/bash
# two GPUs: NCCL gather of device-resident frame records
mkdir -p gpurun_out
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu) > gpurun_out/c8_bench_2gpu.log 2>&1
grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"phases_ms": {[^}]*}\|"e2e": {[^}]*}' gpurun_out/c8_bench_2gpu.log | head; tail -n 5 gpurun_out/c8_bench_2gpu.log | cut -c1-300
