#!/bin/bash
# fc1 alone (stage 2): what bounds the kernel?  Output: gpurun_out/fc1_variants2.jsonl
out=gpurun_out/fc1_variants2.jsonl
: > $out
for m in 18944; do
  for mode in burst; do
    RODC_FC1_BN=256 python profiles/fc1_variants_probe.py $m $mode >> $out 2>gpurun_out/fc1_variants.err
    RODC_FC1_BN=256 RODC_FC1_PROBE_HI_ONLY=1 python profiles/fc1_variants_probe.py $m $mode >> $out 2>>gpurun_out/fc1_variants.err
    RODC_FC1_BN=256 RODC_FC1_KC=4 python profiles/fc1_variants_probe.py $m $mode >> $out 2>>gpurun_out/fc1_variants.err
    RODC_FC1_BN=256 RODC_FC1_KC=4 RODC_FC1_PROBE_HI_ONLY=1 python profiles/fc1_variants_probe.py $m $mode >> $out 2>>gpurun_out/fc1_variants.err
    RODC_FC1_BN=128 RODC_FC1_PROBE_HI_ONLY=1 python profiles/fc1_variants_probe.py $m $mode >> $out 2>>gpurun_out/fc1_variants.err
    RODC_FC1_2CTA=1 python profiles/fc1_variants_probe.py $m $mode >> $out 2>>gpurun_out/fc1_variants.err
    RODC_FC1_2CTA=1 RODC_FC1_KC=2 python profiles/fc1_variants_probe.py $m $mode >> $out 2>>gpurun_out/fc1_variants.err
    RODC_FC1_2CTA=1 RODC_FC1_KC=4 python profiles/fc1_variants_probe.py $m $mode >> $out 2>>gpurun_out/fc1_variants.err
    RODC_FC1_2CTA=1 RODC_FC1_PROBE_HI_ONLY=1 python profiles/fc1_variants_probe.py $m $mode >> $out 2>>gpurun_out/fc1_variants.err
    RODC_FC1_2CTA=1 RODC_FC1_KC=4 RODC_FC1_PROBE_HI_ONLY=1 python profiles/fc1_variants_probe.py $m $mode >> $out 2>>gpurun_out/fc1_variants.err
  done
done
cut -c1-260 $out
