#!/bin/bash
# GPU session r2c: full GPU tests, default bench incl. from-files e2e, kernel timeline of one un-batched frame,
# ncu --set full of the feature kernels with the permuted K order
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -4 gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --from-files > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --config 3 --steps 3 --no-cpu-baseline --from-files > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err; echo "bench c3 rc=$?"
timeout 300 python profiles/ncu_step.py --frames 8 > gpurun_out/r2c_step_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'features_stream' -o gpurun_out/r2c_features -f python profiles/ncu_step.py --frames 8 > gpurun_out/r2c_ncu_features.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches_single_frame.csv python profiles/ncu_step.py --frames 1 > gpurun_out/r2c_ncu_single.log 2>&1; echo "ncu single rc=$?"
ls -la gpurun_out | tail -12
