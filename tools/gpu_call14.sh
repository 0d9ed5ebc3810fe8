#!/bin/bash
# one GPU: the other workloads again after the carry-chain prediction fix (full-size parity gate each)
mkdir -p gpurun_out
for w in nfcb106 mixed nfca424; do
  (time timeout 600 python bench.py --workload $w --steps 2 --warmup 1 --no-e2e --no-wav-set) > gpurun_out/c14_bench_$w.log 2>&1
  echo "== $w"; grep -o '"value": [0-9.]*\|"ms_lanes": [0-9.]*\|"differing": [0-9]*\|"rounds": [0-9]*\|"lane_runs": [0-9]*' gpurun_out/c14_bench_$w.log | tr '\n' ' '; echo
done
