#!/bin/bash
# What round 2 ran on the B200 boxes to produce profiles/r02_*: see the numbered scripts next to this one.
#   tools/gpu_call9a.sh   GPU tests of the long-capture carry exchange + the default bench line (every leg)            1 GPU
#   tools/gpu_call9b.sh   bench lines of the other workloads (nfcb106, mixed, nfca424), exact mode, ncu launch list    1 GPU
#   tools/gpu_call9c.sh   ncu --set full captures: lanes_kernel, screen_kernel, wlanes_kernel, front_kernel            1 GPU
#   tools/gpu_call8.sh    weak scaling at 2 GPUs (torchrun), tools/gpu_call10.sh at 8 GPUs
#   tools/gpu_call11.sh   one continuous capture time-sharded over 2 GPUs with the carry exchange
# and, here (no GPU):  python tools/bench_summary.py ... > profiles/r02_summary.md ;  python tools/ncu_summary.py x.ncu-rep > profiles/r02_x.md ;
#                      python tools/ncu_lines.py x.ncu-rep ;  python tools/sass_summary.py > profiles/r02_sass_summary.md
for s in 9a 9b 9c; do bash tools/gpu_call$s.sh; done
