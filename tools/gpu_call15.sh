#!/bin/bash
# one GPU: compute-sanitizer memcheck over a cut of the GPU tests (batch decode both modes, carry exchange, streaming),
# then the whole GPU suite once more on the final state
mkdir -p gpurun_out
(time timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_long_capture.py -m gpu -q -x \
   -k "test_batch_decode_equals_reference and (NFC-A_106kbps_001 or NFC-V_26kbps_001 or POLL_ABF) or carry_exchange_makes and NFC-A_424kbps_001 or stream") > gpurun_out/c15_sanitizer.log 2>&1
echo "sanitizer rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/c15_sanitizer.log | tail -4
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/c15_gpu_tests.log 2>&1
tail -n 5 gpurun_out/c15_gpu_tests.log | head -2
