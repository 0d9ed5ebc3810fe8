set -x
(time timeout 600 python -m pytest tests -m gpu -q -x) > gpurun_out/c4_gpu_tests.log 2>&1
for v in "2 4 1" "2 4 0" "2 6 1" "2 8 1" "0 4 1" "0 6 1"; do
  set -- $v
  (NFCB200_LANE_TAPS=$1 NFCB200_LANE_BLOCKS=$2 NFCB200_LANE_CG=$3 timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c4_bench_t$1_b$2_g$3.log 2>&1
done
tail -3 gpurun_out/c4_gpu_tests.log; grep -h -o '"value": [0-9.]*\|"ms_lanes": [0-9.]*\|"ms_gather": [0-9.]*' gpurun_out/c4_bench_*.log
