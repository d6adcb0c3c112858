#!/bin/bash
# GPU session r2h: measured upper bound of producer fusion / L2-resident A (profiles/fusion_bound_probe.py)
set -x
mkdir -p gpurun_out
timeout 300 python profiles/fusion_bound_probe.py > gpurun_out/r2h_product.json 2> gpurun_out/r2h_product.err; echo "rc=$?"
RODC_PROBE_A_WRAP=256 timeout 300 python profiles/fusion_bound_probe.py --probe > gpurun_out/r2h_probe_wrap256.json 2> gpurun_out/r2h_probe_wrap256.err; echo "rc=$?"
RODC_PROBE_A_WRAP=0 timeout 300 python profiles/fusion_bound_probe.py --probe > gpurun_out/r2h_probe_lib_no_wrap.json 2> gpurun_out/r2h_probe_lib_no_wrap.err; echo "rc=$?"
timeout 300 python profiles/fusion_bound_probe.py > gpurun_out/r2h_product_again.json 2>> gpurun_out/r2h_product.err; echo "rc=$?"
cat gpurun_out/r2h_*.json
