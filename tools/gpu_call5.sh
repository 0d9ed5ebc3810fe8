#!/bin/bash
mkdir -p gpurun_out
(time timeout 300 python tools/gpu_probe.py) > gpurun_out/c5_probe_fast.log 2>&1
(time PROBE_EXACT=1 timeout 300 python tools/gpu_probe.py) > gpurun_out/c5_probe_exact.log 2>&1
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/c5_gpu_tests.log 2>&1
(time timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c5_bench_full.log 2>&1
grep -c OK gpurun_out/c5_probe_fast.log gpurun_out/c5_probe_exact.log
tail -n 3 gpurun_out/c5_gpu_tests.log
grep -o '"value": [0-9.]*\|"phases_ms": {[^}]*}\|"frames_digest": "[0-9a-f]*"' gpurun_out/c5_bench_full.log
