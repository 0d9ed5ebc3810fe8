#!/bin/bash
# one GPU: the lane kernel with the straggler pass off and on, split times (NFCB200_TRACE)
mkdir -p gpurun_out
for m in 0 4096; do
  (NFCB200_TRACE=1 NFCB200_STRAGGLER=$m timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c17_bench_m$m.log 2>&1
  echo "== margin $m"; grep -o '"ms_lanes": [0-9.]*' gpurun_out/c17_bench_m$m.log; grep "thread lanes\|straggler\|round 1" gpurun_out/c17_bench_m$m.log | tail -3
done
