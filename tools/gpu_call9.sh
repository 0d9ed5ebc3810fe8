set -x
(time NFCB200_SCREEN_DB=0 timeout 300 python -m pytest tests -m gpu -q -x) > gpurun_out/c9_gpu_tests_default.log 2>&1
(time NFCB200_SCREEN_DB=1 timeout 300 python -m pytest tests -m gpu -q -x) > gpurun_out/c9_gpu_tests_screen_db.log 2>&1
(NFCB200_SCREEN_DB=0 timeout 150 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c9_bench_default.log 2>&1
(NFCB200_SCREEN_DB=1 timeout 150 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c9_bench_screen_db.log 2>&1
grep -h -E "passed|failed" gpurun_out/c9_gpu_tests_default.log gpurun_out/c9_gpu_tests_screen_db.log; grep -h -o '"value": [0-9.]*\|"ms_screen": [0-9.]*\|"frames_digest": "[0-9a-f]*"' gpurun_out/c9_bench_*.log
