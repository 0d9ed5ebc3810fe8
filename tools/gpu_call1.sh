set -x
(nproc; cat /sys/fs/cgroup/cpu.max; python -c "import os;print(os.cpu_count(), len(os.sched_getaffinity(0)))"; free -g | head -2; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv) > gpurun_out/c1_sysinfo.txt 2>&1

(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/c1_gpu_tests.log 2>&1
(time timeout 600 python bench.py) > gpurun_out/c1_bench_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c1_launches.csv python bench.py --quick --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/c1_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 1 -c 1 -f -o gpurun_out/c1_screen python bench.py --quick --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/c1_ncu_screen.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lanes_kernel -s 1 -c 1 -f -o gpurun_out/c1_lanes python bench.py --quick --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/c1_ncu_lanes.log 2>&1
tail -3 gpurun_out/c1_gpu_tests.log; tail -2 gpurun_out/c1_bench_full.log
