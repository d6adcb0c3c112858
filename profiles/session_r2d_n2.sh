#!/bin/bash
# GPU session r2d (2 GPUs): N ranks == 1 rank over NCCL (pytest), bench at N = 2 (config 2 and config 3)
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2d_gpus.txt
timeout 900 python -m pytest tests/test_gpu_files_and_callers.py -m gpu -q -k "nccl" > gpurun_out/r2d_nccl_test.log 2>&1; echo "nccl test rc=$?" >> gpurun_out/r2d_nccl_test.log
tail -3 gpurun_out/r2d_nccl_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; echo "bench n2 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 3 --steps 3 --no-kernel-table > gpurun_out/r2d_bench_c3_n2.json 2> gpurun_out/r2d_bench_c3_n2.err; echo "bench c3 n2 rc=$?"
ls -la gpurun_out | tail -8
