#!/bin/bash
# two GPUs: one continuous capture, time-sharded, carry exchange rank to rank, parity gate against the uncut decode
mkdir -p gpurun_out
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/long_capture_bench.py --samples 200000000) > gpurun_out/c11_long_capture_2gpu.log 2>&1
tail -n 6 gpurun_out/c11_long_capture_2gpu.log | cut -c1-900
