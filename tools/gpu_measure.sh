#!/bin/bash
# What round 1 ran on the B200 box (one GPU) to produce profiles/: GPU parity suite, the bench line, the ncu launch list and
# the two full ncu captures.  Usage on the box: bash tools/gpu_measure.sh   (outputs under gpurun_out/)
set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/gpu_tests.log 2>&1
(time timeout 900 python bench.py) > gpurun_out/bench_full.log 2>&1
(time timeout 300 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/bench_reference.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'screen_kernel|segment_|lanes_kernel|chain_|lane_' -c 200 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 1 -c 1 -f -o gpurun_out/screen \
    python bench.py --quick --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_screen.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lanes_kernel -c 1 -f -o gpurun_out/lanes \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_lanes.log 2>&1
# then, here: python tools/ncu_summary.py gpurun_out/screen.ncu-rep "title" > profiles/rNN_screen_kernel.md ; python tools/ncu_lines.py gpurun_out/lanes.ncu-rep
