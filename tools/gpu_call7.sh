set -x
(time timeout 900 python bench.py) > gpurun_out/c7_bench_full.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'screen_kernel|segment_|lanes_kernel|chain_kernel|lane_' -c 200 --csv --log-file gpurun_out/c7_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/c7_launches_bench.log 2>&1
tail -2 gpurun_out/c7_bench_full.log | cut -c1-600
