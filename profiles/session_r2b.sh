#!/bin/bash
# GPU session r2b: full GPU tests (latency path, staged re-crop, file-loaded cascades, callers), default bench,
# config-3 bench, targeted ncu of the staged re-crop and of one un-batched frame
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --config 3 --steps 3 --no-cpu-baseline > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err; echo "bench c3 rc=$?"
timeout 300 python profiles/ncu_step.py --frames 8 > gpurun_out/r2b_step_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'recrop|compact_gather|group_union|pyr_' -o gpurun_out/r2b_small -f python profiles/ncu_step.py --frames 8 > gpurun_out/r2b_ncu_small.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -12
