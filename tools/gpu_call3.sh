#!/bin/bash
mkdir -p gpurun_out
(time timeout 300 python tools/gpu_probe.py) > gpurun_out/c3_probe.log 2>&1
(time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x) > gpurun_out/c3_gpu_tests.log 2>&1
(time timeout 300 python bench.py --quick --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c3_bench_quick.log 2>&1
(time timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c3_bench_full.log 2>&1
tail -n 4 gpurun_out/c3_probe.log gpurun_out/c3_gpu_tests.log
grep -o '"value": [0-9.]*\|"phases_ms": {[^}]*}\|"frames_digest": "[0-9a-f]*"' gpurun_out/c3_bench_quick.log gpurun_out/c3_bench_full.log
