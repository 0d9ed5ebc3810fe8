#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wlanes_kernel -c 1 -f -o gpurun_out/c2_wlanes \
    python bench.py --quick --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/c2_ncu_wlanes.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:front_kernel -c 1 -f -o gpurun_out/c2_front \
    python bench.py --quick --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/c2_ncu_front.log 2>&1
tail -n 3 gpurun_out/c2_ncu_wlanes.log gpurun_out/c2_ncu_front.log
