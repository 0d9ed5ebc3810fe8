#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^lanes_kernel -c 1 -f -o gpurun_out/c9_lanes \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_lanes.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 1 -c 1 -f -o gpurun_out/c9_screen \
    python bench.py --steps 2 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_screen.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wlanes_kernel -c 1 -f -o gpurun_out/c9_wlanes \
    python bench.py --exact --quick --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_wlanes.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:front_kernel -c 1 -f -o gpurun_out/c9_front \
    python bench.py --exact --quick --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_front.log 2>&1
ls -la gpurun_out/*.ncu-rep
