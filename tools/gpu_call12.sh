#!/bin/bash
# one GPU: the batches ranks 2 and 3 of a multi-GPU run decode (seed + 1000 * rank), alone
mkdir -p gpurun_out
for r in 0 2; do
  (NFCB200_TRACE=1 timeout 600 python bench.py --seed $((2024 + 1000 * r)) --steps 1 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c12_bench_seed_r$r.log 2>&1
  grep -o '"ms_per_step": [0-9.]*\|"decode": {[^}]*}\|"phases_ms": {[^}]*}' gpurun_out/c12_bench_seed_r$r.log
  grep "longest\|runs by\|round" gpurun_out/c12_bench_seed_r$r.log | tail -3 | cut -c1-1500
done
