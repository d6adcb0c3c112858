#!/bin/bash
# GPU session r2p: last session of round 2 -- full GPU tests, smoke, the bench lines of every BASELINE config at one GPU
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2p_tests.log
tail -3 gpurun_out/r2p_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --from-files > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --config 3 --steps 3 --no-cpu-baseline --from-files > gpurun_out/r2p_bench_c3.json 2> gpurun_out/r2p_bench_c3.err; echo "bench c3 rc=$?"
timeout 900 python bench.py --config 4 --steps 3 --no-cpu-baseline > gpurun_out/r2p_bench_c4.json 2> gpurun_out/r2p_bench_c4.err; echo "bench c4 rc=$?"
timeout 900 python bench.py --config 5 --steps 3 --no-cpu-baseline > gpurun_out/r2p_bench_c5.json 2> gpurun_out/r2p_bench_c5.err; echo "bench c5 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-table > gpurun_out/r2p_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
