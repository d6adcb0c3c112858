#!/bin/bash
# GPU session r2i: full GPU tests + smoke + default bench after the dp2a re-crop and the host-path trims
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log
tail -3 gpurun_out/r2i_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --config 3 --steps 3 --no-cpu-baseline --no-kernel-table > gpurun_out/r2i_bench_c3.json 2> gpurun_out/r2i_bench_c3.err; echo "bench c3 rc=$?"
