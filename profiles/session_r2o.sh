#!/bin/bash
# GPU session r2o: final verification of the round-2 tree: full GPU tests, smoke, default bench, config 3, reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2o_tests.log
tail -3 gpurun_out/r2o_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --from-files > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --config 3 --steps 3 --no-cpu-baseline --no-kernel-table --from-files > gpurun_out/r2o_bench_c3.json 2> gpurun_out/r2o_bench_c3.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2o_bench_reference.json 2> gpurun_out/r2o_bench_reference.err; echo "bench ref rc=$?"
