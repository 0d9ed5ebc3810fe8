#!/bin/bash
# one GPU: the straggler tests and a cut of the parity tests on the final binary
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "straggler or (test_batch_decode_equals_reference and 106kbps_001)") > gpurun_out/c20_gpu_tests.log 2>&1
tail -n 5 gpurun_out/c20_gpu_tests.log | head -2
