#!/bin/bash
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_tap.py -m gpu -q) > gpurun_out/c7_tap.log 2>&1
(time NFCB200_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c7_bench_trace.log 2>&1
(time timeout 900 python bench.py --exact --steps 2 --warmup 1 --no-e2e --no-wav-set) > gpurun_out/c7_bench_exact.log 2>&1
tail -n 4 gpurun_out/c7_tap.log
grep "\[nfcb200\]" gpurun_out/c7_bench_trace.log | tail -14
grep -o '"value": [0-9.]*\|"phases_ms": {[^}]*}\|"frames_digest": "[0-9a-f]*"\|"full_parity": {[^}]*}[^}]*}' gpurun_out/c7_bench_trace.log gpurun_out/c7_bench_exact.log
