#!/bin/bash
# GPU session r2e (8 GPUs): bench at N = 8 and N = 4 (config 2), config 3 (strong scaling) at N = 8 and 4
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus.txt
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 --no-kernel-table > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err; echo "bench n$N rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --config 3 --steps 3 --no-kernel-table > gpurun_out/r2e_bench_c3_n$N.json 2> gpurun_out/r2e_bench_c3_n$N.err; echo "bench c3 n$N rc=$?"
done
ls -la gpurun_out | tail -8
