#!/bin/bash
# GPU session r2r: throughput-form stage-0 window copy: full GPU tests + smoke + configs 2 / 3 / 4 / 5
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2r_tests.log
tail -3 gpurun_out/r2r_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --config 3 --steps 3 --no-cpu-baseline > gpurun_out/r2r_bench_c3.json 2> gpurun_out/r2r_bench_c3.err; echo "bench c3 rc=$?"
timeout 900 python bench.py --config 4 --steps 3 --no-cpu-baseline > gpurun_out/r2r_bench_c4.json 2> gpurun_out/r2r_bench_c4.err; echo "bench c4 rc=$?"
timeout 900 python bench.py --config 5 --steps 3 --no-cpu-baseline > gpurun_out/r2r_bench_c5.json 2> gpurun_out/r2r_bench_c5.err; echo "bench c5 rc=$?"
