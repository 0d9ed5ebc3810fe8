#!/bin/bash
# one GPU: the default lane kernel (straggler hand-over not compiled into it) once more
mkdir -p gpurun_out
(NFCB200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu) > gpurun_out/c18_bench_default.log 2>&1
grep -o '"ms_lanes": [0-9.]*\|"value": [0-9.]*' gpurun_out/c18_bench_default.log | head -3; grep "round 1" gpurun_out/c18_bench_default.log | tail -2
