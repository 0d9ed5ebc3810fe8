#!/bin/bash
# one GPU, the final state of round 2: the whole GPU test suite, the default bench line (every leg), the reference arm,
# the other workloads with their full-size parity gate, exact mode at full size
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/c13_gpu_tests.log 2>&1
tail -n 2 gpurun_out/c13_gpu_tests.log | head -1
(time timeout 900 python bench.py) > gpurun_out/c13_bench_default.log 2>&1
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/c13_bench_reference.log 2>&1
for w in nfcb106 mixed nfca424; do
  (time timeout 600 python bench.py --workload $w --steps 2 --warmup 1 --no-e2e --no-wav-set) > gpurun_out/c13_bench_$w.log 2>&1
done
(time timeout 900 python bench.py --exact --steps 2 --warmup 1 --no-e2e --no-wav-set) > gpurun_out/c13_bench_exact.log 2>&1
for f in default reference nfcb106 mixed nfca424 exact; do
  echo "== $f"; grep -o '"value": [0-9.]*\|"ms_lanes": [0-9.]*\|"differing": [0-9]*\|"rounds": [0-9]*\|"lane_runs": [0-9]*' gpurun_out/c13_bench_$f.log | tr '\n' ' '; echo; done
