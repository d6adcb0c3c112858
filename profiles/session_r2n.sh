#!/bin/bash
# GPU session r2n: occupancy experiment of the 24 / 48-px feature kernels (three CTAs per SM with spills vs two without)
set -x
mkdir -p gpurun_out
timeout 300 python profiles/fusion_bound_probe.py --probe > gpurun_out/r2n_occ2.json 2> gpurun_out/r2n_occ2.err; echo "rc=$?"
RODC_PROBE_FEAT_OCC3=1 timeout 300 python profiles/fusion_bound_probe.py --probe > gpurun_out/r2n_occ3.json 2> gpurun_out/r2n_occ3.err; echo "rc=$?"
timeout 300 python profiles/fusion_bound_probe.py --probe > gpurun_out/r2n_occ2_again.json 2>> gpurun_out/r2n_occ2.err; echo "rc=$?"
RODC_PROBE_FEAT_OCC3=1 timeout 300 python profiles/fusion_bound_probe.py --probe > gpurun_out/r2n_occ3_again.json 2>> gpurun_out/r2n_occ3.err; echo "rc=$?"
cat gpurun_out/r2n_*.json
