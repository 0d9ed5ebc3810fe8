#!/bin/bash
# one GPU: phase counters of exact mode on the final binary (64 x 2e6 batch)
mkdir -p gpurun_out
(NFCB200_TRACE=1 timeout 100 python bench.py --exact --quick --steps 1 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c21_bench_exact_quick.log 2>&1
grep "nfcb200\] lanes " gpurun_out/c21_bench_exact_quick.log | tail -9; grep -o '"ms_lanes": [0-9.]*' gpurun_out/c21_bench_exact_quick.log
