#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/c6_gpu_tests.log 2>&1
(time timeout 900 python bench.py --steps 3 --warmup 1 --no-e2e) > gpurun_out/c6_bench_full.log 2>&1
tail -n 4 gpurun_out/c6_gpu_tests.log
grep -o '"value": [0-9.]*\|"phases_ms": {[^}]*}\|"frames_digest": "[0-9a-f]*"\|"full_parity": {[^}]*}[^}]*}\|"cpu_baseline": {[^}]*}' gpurun_out/c6_bench_full.log
