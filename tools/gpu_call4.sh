#!/bin/bash
mkdir -p gpurun_out
(time NFCB200_TRACE=1 timeout 300 python tools/gpu_probe.py) > gpurun_out/c4_probe.log 2>&1
(time NFCB200_TRACE=1 timeout 300 python bench.py --quick --steps 1 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c4_bench_quick.log 2>&1
(time NFCB200_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c4_bench_full.log 2>&1
grep -c OK gpurun_out/c4_probe.log
grep "lanes \|lanes + chain\|front" gpurun_out/c4_bench_full.log | tail -12
grep -o '"value": [0-9.]*\|"phases_ms": {[^}]*}\|"frames_digest": "[0-9a-f]*"' gpurun_out/c4_bench_quick.log gpurun_out/c4_bench_full.log
