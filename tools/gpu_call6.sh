set -x
(nproc; cat /sys/fs/cgroup/cpu.max; cat /sys/fs/cgroup/memory.max; cat /sys/fs/cgroup/memory.current; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv) > gpurun_out/c6_sysinfo.txt 2>&1
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/c6_gpu_tests.log 2>&1
(NFCB200_HALO_SHORT=0 timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu) > gpurun_out/c6_bench_longhalo.log 2>&1
(time timeout 900 python bench.py) > gpurun_out/c6_bench_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c6_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/c6_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 1 -c 1 -f -o gpurun_out/c6_screen python bench.py --quick --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/c6_ncu_screen.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lanes_kernel -c 1 -f -o gpurun_out/c6_lanes python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > gpurun_out/c6_ncu_lanes.log 2>&1
(time timeout 300 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/c6_bench_reference.log 2>&1
tail -3 gpurun_out/c6_gpu_tests.log; tail -2 gpurun_out/c6_bench_full.log | cut -c1-400
