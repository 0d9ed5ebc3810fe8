#!/bin/bash
# What round 2 ran on the B200 boxes to produce profiles/r02_*: the numbered scripts next to this one, in this order.
#   tools/gpu_call9a.sh   GPU tests of the long-capture carry exchange + the default bench line (every leg)            1 GPU
#   tools/gpu_call9b.sh   bench lines of the other workloads, exact mode (quick, with phase counters), ncu launch list  1 GPU
#   tools/gpu_call9c.sh   ncu --set full captures: lanes_kernel, screen_kernel, wlanes_kernel, front_kernel            1 GPU
#   tools/gpu_call8.sh    weak scaling at 2 GPUs, tools/gpu_call10b.sh at 4 (tools/gpu_call10.sh: 8), per-rank phases
#   tools/gpu_call11.sh   one continuous capture time-sharded over 2 GPUs with the carry exchange (NCCL)
#   tools/gpu_call12.sh   the batches ranks 0 / 2 of a multi-GPU run decode, alone: run-length histogram, longest lane
#   tools/gpu_call13.sh   final state: GPU suite, default line, reference arm, other workloads, exact mode at full size
#   tools/gpu_call14.sh   the other workloads after the carry-chain prediction fix
#   tools/gpu_call15.sh   compute-sanitizer memcheck over a cut of the GPU tests + the GPU suite
#   tools/gpu_call16.sh .. 19   the straggler hand-over experiment (suite, seeds, split times, phase counters)
# and, here (no GPU):  python tools/bench_summary.py title logs... ;  python tools/ncu_summary.py x.ncu-rep ;  python tools/ncu_lines.py x.ncu-rep ;
#                      python tools/sass_summary.py > profiles/r02_sass_summary.md
for s in 13 15; do bash tools/gpu_call$s.sh; done
