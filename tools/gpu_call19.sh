#!/bin/bash
# one GPU: phase counters of the warp lane that decodes the straggler (NFCB200_STRAGGLER=4096)
mkdir -p gpurun_out
(NFCB200_TRACE=1 NFCB200_STRAGGLER=4096 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c19_bench_m4096.log 2>&1
grep "nfcb200\] lanes \|straggler\|thread lanes" gpurun_out/c19_bench_m4096.log | tail -12
