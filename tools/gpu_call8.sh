set -x
(time NFCB200_CHAIN_WARP=0 timeout 400 python -m pytest tests -m gpu -q -x) > gpurun_out/c8_gpu_tests_scalar_chain.log 2>&1
(time NFCB200_CHAIN_WARP=1 timeout 400 python -m pytest tests -m gpu -q -x) > gpurun_out/c8_gpu_tests_warp_chain.log 2>&1
(NFCB200_CHAIN_WARP=0 timeout 200 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c8_bench_scalar_chain.log 2>&1
(NFCB200_CHAIN_WARP=1 timeout 200 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c8_bench_warp_chain.log 2>&1
tail -2 gpurun_out/c8_gpu_tests_scalar_chain.log gpurun_out/c8_gpu_tests_warp_chain.log; grep -h -o '"value": [0-9.]*\|"ms_lanes": [0-9.]*\|"frames_digest": "[0-9a-f]*"' gpurun_out/c8_bench_*.log
