#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_long_capture.py -m gpu -q) > gpurun_out/c9_long_capture_tests.log 2>&1
tail -n 3 gpurun_out/c9_long_capture_tests.log
for w in nfcb106 mixed nfca424; do
  (time timeout 600 python bench.py --workload $w --steps 2 --warmup 1 --no-e2e --no-wav-set) > gpurun_out/c9_bench_$w.log 2>&1
done
(time NFCB200_TRACE=1 timeout 600 python bench.py --exact --quick --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c9_bench_exact_quick.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c9_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/c9_launches_bench.log 2>&1
for f in gpurun_out/c9_bench_nfcb106.log gpurun_out/c9_bench_mixed.log gpurun_out/c9_bench_nfca424.log gpurun_out/c9_bench_exact_quick.log; do
  echo "== $f"; grep -o '"value": [0-9.]*\|"ms_lanes": [0-9.]*\|"differing": [0-9]*\|"frames": [0-9]*,' $f | tr '\n' ' '; echo; done
