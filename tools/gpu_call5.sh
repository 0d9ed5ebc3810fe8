set -x
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1) > gpurun_out/c5_bench_2gpu.log 2>&1
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1) > gpurun_out/c5_ref_2gpu.log 2>&1
tail -5 gpurun_out/c5_bench_2gpu.log | cut -c1-1500; tail -3 gpurun_out/c5_ref_2gpu.log | cut -c1-600
