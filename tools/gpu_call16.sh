#!/bin/bash
# one GPU: the straggler pass -- GPU suite, the default line, the two batches whose lane kernel ended with a straggler
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/c16_gpu_tests.log 2>&1
tail -n 5 gpurun_out/c16_gpu_tests.log | head -2
for r in 0 2 3; do
  (NFCB200_TRACE=1 timeout 600 python bench.py --seed $((2024 + 1000 * r)) --steps 2 --warmup 1 --no-e2e --no-wav-set) > gpurun_out/c16_bench_seed_r$r.log 2>&1
  echo "== seed r$r"; grep -o '"value": [0-9.]*\|"ms_lanes": [0-9.]*\|"differing": [0-9]*\|"rounds": [0-9]*' gpurun_out/c16_bench_seed_r$r.log | tr '\n' ' '; echo
  grep "straggler\|longest" gpurun_out/c16_bench_seed_r$r.log | tail -2 | cut -c1-200
done
(time timeout 900 python bench.py) > gpurun_out/c16_bench_default.log 2>&1
echo "== default"; grep -o '"value": [0-9.]*\|"ms_lanes": [0-9.]*\|"differing": [0-9]*' gpurun_out/c16_bench_default.log | tr '\n' ' '; echo
