#!/bin/bash
# GPU session r2e (8 GPUs): bench at N = 8: config 2 (weak), config 3 (strong, 2845 frames), config 4 (1080p stream on 8 GPUs)
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus.txt; nproc >> gpurun_out/r2e_gpus.txt
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus $N --steps 10 --warmup 3 --no-kernel-table > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err; echo "bench n$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus $N --config 3 --steps 5 --no-kernel-table > gpurun_out/r2e_bench_c3_n$N.json 2> gpurun_out/r2e_bench_c3_n$N.err; echo "bench c3 n$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus $N --config 4 --steps 3 --no-kernel-table > gpurun_out/r2e_bench_c4_n$N.json 2> gpurun_out/r2e_bench_c4_n$N.err; echo "bench c4 n$N rc=$?"
ls -la gpurun_out | tail -8
