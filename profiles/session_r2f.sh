#!/bin/bash
# GPU session r2f: the tests touched since r2c, config-3 e2e after the host-path changes, config 4 / 5 benches
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q -k "row_output or pipeline or files or sources or error_conventions or config3 or sharded" > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log
tail -3 gpurun_out/r2f_tests.log
timeout 900 python bench.py --config 3 --steps 3 --no-cpu-baseline --no-kernel-table > gpurun_out/r2f_bench_c3.json 2> gpurun_out/r2f_bench_c3.err; echo "bench c3 rc=$?"
timeout 900 python bench.py --config 4 --steps 3 > gpurun_out/r2f_bench_c4.json 2> gpurun_out/r2f_bench_c4.err; echo "bench c4 rc=$?"
timeout 900 python bench.py --config 5 --steps 3 > gpurun_out/r2f_bench_c5.json 2> gpurun_out/r2f_bench_c5.err; echo "bench c5 rc=$?"
ls -la gpurun_out | tail -6
