set -x
(time timeout 600 python -m pytest tests -m gpu -q -x) > gpurun_out/c3_gpu_tests.log 2>&1
for v in "2 4" "0 4" "2 3"; do
  set -- $v
  (NFCB200_LANE_TAPS=$1 NFCB200_LANE_BLOCKS=$2 timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c3_bench_t$1_b$2.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lanes_kernel -c 1 -f -o gpurun_out/c3_lanes python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > gpurun_out/c3_ncu_lanes.log 2>&1
(time timeout 900 python bench.py) > gpurun_out/c3_bench_full.log 2>&1
tail -3 gpurun_out/c3_gpu_tests.log; grep -h -o '"value": [0-9.]*\|"ms_lanes": [0-9.]*\|"ms_screen": [0-9.]*' gpurun_out/c3_bench_*.log
