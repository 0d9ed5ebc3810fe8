#!/bin/bash
# one GPU: ncu --set full of the four main kernels; the reports stay on the box (> 64 MiB together), their summaries
# (tools/ncu_summary.py, tools/ncu_lines.py) and the two lane-kernel reports come back
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^lanes_kernel -c 1 -f -o /tmp/ncu/r02_lanes \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_lanes.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 1 -c 1 -f -o /tmp/ncu/r02_screen \
    python bench.py --steps 2 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_screen.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wlanes_kernel -c 1 -f -o /tmp/ncu/r02_wlanes \
    python bench.py --exact --quick --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_wlanes.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:front_kernel -c 1 -f -o /tmp/ncu/r02_front \
    python bench.py --exact --quick --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/c9_ncu_front.log 2>&1
for k in lanes screen wlanes front; do
  python tools/ncu_summary.py /tmp/ncu/r02_$k.ncu-rep "round 2: ${k} kernel, ncu --set full" > gpurun_out/r02_${k}_kernel.md 2>gpurun_out/r02_${k}_summary.err
  python tools/ncu_lines.py /tmp/ncu/r02_$k.ncu-rep 40 >> gpurun_out/r02_${k}_kernel.md 2>>gpurun_out/r02_${k}_summary.err
  ncu -i /tmp/ncu/r02_$k.ncu-rep --page raw --csv > gpurun_out/r02_${k}_raw.csv 2>/dev/null
done
ls -la /tmp/ncu
cp /tmp/ncu/r02_wlanes.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out
head -c 1500 gpurun_out/r02_lanes_kernel.md
