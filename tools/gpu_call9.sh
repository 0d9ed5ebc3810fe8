#!/bin/bash
# round 2 measurement set (one GPU): default bench with every leg, the other workloads, launch list, ncu captures
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_long_capture.py -m gpu -q -x) > gpurun_out/c9_long_capture_tests.log 2>&1
(time timeout 900 python bench.py) > gpurun_out/c9_bench_default.log 2>&1
for w in nfcb106 mixed nfca424; do
  (time timeout 600 python bench.py --workload $w --steps 2 --warmup 1 --no-e2e --no-wav-set) > gpurun_out/c9_bench_$w.log 2>&1
done
(time NFCB200_TRACE=1 timeout 600 python bench.py --exact --quick --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c9_bench_exact_quick.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c9_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/c9_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^lanes_kernel -c 1 -f -o gpurun_out/c9_lanes \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_lanes.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 1 -c 1 -f -o gpurun_out/c9_screen \
    python bench.py --steps 2 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_screen.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wlanes_kernel -c 1 -f -o gpurun_out/c9_wlanes \
    python bench.py --exact --quick --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_wlanes.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:front_kernel -c 1 -f -o gpurun_out/c9_front \
    python bench.py --exact --quick --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_front.log 2>&1
for f in gpurun_out/c9_bench_default.log gpurun_out/c9_bench_nfcb106.log gpurun_out/c9_bench_mixed.log gpurun_out/c9_bench_nfca424.log gpurun_out/c9_bench_exact_quick.log; do
  echo "== $f"; grep -o '"value": [0-9.]*\|"ms_lanes": [0-9.]*\|"differing": [0-9]*\|"frames": [0-9]*,' $f | tr '\n' ' '; echo; done
tail -n 4 gpurun_out/c9_long_capture_tests.log
