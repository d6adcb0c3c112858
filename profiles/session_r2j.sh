#!/bin/bash
# GPU session r2j: host profile of the config-3 e2e path; pyramid tests + ncu of the TMA-staged h-pass
set -x
mkdir -p gpurun_out
timeout 600 python profiles/e2e_host_profile.py 1300 > gpurun_out/r2j_e2e_profile.txt 2>&1; echo "rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q -k "pyramid or windows or config or latency or pipeline or cascade_app" > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log
tail -3 gpurun_out/r2j_tests.log
timeout 300 python profiles/ncu_step.py --frames 8 > gpurun_out/r2j_step_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'pyr_|recrop|source_windows' -o gpurun_out/r2j_small -f python profiles/ncu_step.py --frames 8 > gpurun_out/r2j_ncu_small.log 2>&1; echo "ncu rc=$?"
