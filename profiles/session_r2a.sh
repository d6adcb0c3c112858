#!/bin/bash
# GPU session r2a: tests, default bench, config-3 bench, ncu --set full of one step, ncu launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --config 3 --frames 560 --steps 3 --no-cpu-baseline > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; echo "bench c3 rc=$?"
timeout 300 python profiles/ncu_step.py --frames 8 > gpurun_out/r2a_step_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r2a_step -f python profiles/ncu_step.py --frames 8 > gpurun_out/r2a_ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-table > gpurun_out/r2a_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out | tail -20
