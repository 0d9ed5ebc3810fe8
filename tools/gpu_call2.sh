set -x
(time timeout 600 python -m pytest tests -m gpu -q) > gpurun_out/c2_gpu_tests.log 2>&1
(time timeout 300 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_parity.py -q -x -k "synthetic and nfca106") > gpurun_out/c2_sanitizer.log 2>&1
for v in "2 4" "0 4" "1 4" "2 3" "1 3" "2 2"; do
  set -- $v
  (NFCB200_LANE_TAPS=$1 NFCB200_LANE_BLOCKS=$2 timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c2_bench_t$1_b$2.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 1 -c 1 -f -o gpurun_out/c2_screen python bench.py --quick --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/c2_ncu_screen.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lanes_kernel -c 1 -f -o gpurun_out/c2_lanes python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > gpurun_out/c2_ncu_lanes.log 2>&1
tail -3 gpurun_out/c2_gpu_tests.log; grep -h -o '"value": [0-9.]*\|"ms_lanes": [0-9.]*\|"ms_screen": [0-9.]*' gpurun_out/c2_bench_*.log
