#!/bin/bash
# round 2, GPU call 1: the warp-lane pipeline on hardware for the first time
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.log 2>&1
(time timeout 300 python tools/gpu_probe.py) > gpurun_out/c1_probe.log 2>&1
(time timeout 600 python -m pytest tests -m gpu -q -x) > gpurun_out/c1_gpu_tests.log 2>&1
(time timeout 300 python bench.py --quick --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c1_bench_quick.log 2>&1
(time timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c1_bench_full.log 2>&1
tail -3 gpurun_out/c1_probe.log gpurun_out/c1_gpu_tests.log gpurun_out/c1_bench_quick.log gpurun_out/c1_bench_full.log
