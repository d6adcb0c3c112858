set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t58.log 2>&1; tail -3 gpurun_out/t58.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1s.log 2>&1; tail -2 gpurun_out/smoke_r1s.log
timeout 400 python profiles/config_throughput.py > gpurun_out/configs_r1s.json 2>gpurun_out/configs_r1s.err; tail -3 gpurun_out/configs_r1s.err
timeout 400 python bench.py > gpurun_out/bench_r1s.log 2>gpurun_out/bench_r1s.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1s.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1s.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_r1s.csv > gpurun_out/launches_r1s_summary.txt; head -12 gpurun_out/launches_r1s_summary.txt
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1s.log 2>&1; tail -1 gpurun_out/bench_ref_r1s.log | cut -c1-400
