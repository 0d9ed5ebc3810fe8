#!/bin/bash
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/c9_mem.log; cat /sys/fs/cgroup/memory.max >> gpurun_out/c9_mem.log 2>&1; nproc >> gpurun_out/c9_mem.log; cat /sys/fs/cgroup/cpu.max >> gpurun_out/c9_mem.log 2>&1
(time timeout 600 python -m pytest tests/test_long_capture.py -m gpu -q -x) > gpurun_out/c9_long_capture_tests.log 2>&1
tail -n 4 gpurun_out/c9_long_capture_tests.log
(time timeout 900 python bench.py) > gpurun_out/c9_bench_default.log 2>&1
grep -o '"value": [0-9.]*\|"ms_lanes": [0-9.]*\|"differing": [0-9]*\|"frames": [0-9]*,' gpurun_out/c9_bench_default.log | tr '\n' ' '; cat gpurun_out/c9_mem.log
