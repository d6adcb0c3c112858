#!/bin/bash
# GPU session r2g (4 GPUs): bench at N = 4: config 2 (weak), config 3 (strong)
set -x
mkdir -p gpurun_out
N=4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus $N --steps 10 --warmup 3 --no-kernel-table > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err; echo "bench n$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --config 3 --steps 5 --no-kernel-table > gpurun_out/r2g_bench_c3_n$N.json 2> gpurun_out/r2g_bench_c3_n$N.err; echo "bench c3 n$N rc=$?"
ls -la gpurun_out | tail -6
