#!/bin/bash
# margin error of the tensor fc1 vs the fp64 oracle for the drain interval KC (k-blocks accumulated inside the tensor core)
out=gpurun_out/fc1_precision.txt
: > $out
for cfg in "RODC_FC1_KC=1" "RODC_FC1_KC=2" "RODC_FC1_KC=4" "RODC_FC1_2CTA=1 RODC_FC1_KC=1" "RODC_FC1_2CTA=1 RODC_FC1_KC=2" "RODC_FC1_2CTA=1 RODC_FC1_KC=3" "RODC_FC1_2CTA=1 RODC_FC1_KC=4"; do
  echo "== $cfg" >> $out
  env $cfg python -m pytest tests/test_gpu_kernels.py -q -s -k "stage_forward_matches and tensor" 2>&1 | grep -oE "stage [0-9] tensor.*|[0-9]+ passed.*|[0-9]+ failed.*" >> $out
done
cat $out
