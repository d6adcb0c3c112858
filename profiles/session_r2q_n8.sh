#!/bin/bash
# GPU session r2q (8 GPUs): config 2 and config 3 at N = 8 on the final tree
set -x
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29628 bench.py --gpus $N --steps 10 --warmup 3 --no-kernel-table > gpurun_out/r2q_bench_n$N.json 2> gpurun_out/r2q_bench_n$N.err; echo "bench n$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29638 bench.py --gpus $N --config 3 --steps 5 --no-kernel-table > gpurun_out/r2q_bench_c3_n$N.json 2> gpurun_out/r2q_bench_c3_n$N.err; echo "bench c3 n$N rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29648 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r2q_bench_reference_n$N.json 2> gpurun_out/r2q_bench_reference_n$N.err; echo "bench ref n$N rc=$?"
