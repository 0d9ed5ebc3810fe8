#!/bin/bash
# eight GPUs: weak scaling of the batch decode (1024 streams per GPU) with the device-record frame gather
mkdir -p gpurun_out
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu --no-e2e) > gpurun_out/c10_bench_8gpu.log 2>&1
grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"phases_ms": {[^}]*}' gpurun_out/c10_bench_8gpu.log | head; tail -n 3 gpurun_out/c10_bench_8gpu.log | cut -c1-300
