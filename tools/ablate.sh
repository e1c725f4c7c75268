#!/bin/bash
# Timing experiment: which role bounds the conv kernels.  SSNB_ABLATE bits: 1 no epilogue stores, 4 empty epilogue,
# 8 no MMAs (both generations); 2 no bias loads, 16 no weight TMA loads, 32 no activation TMA loads (first generation,
# SSNB_V2=0).  Results are garbage by construction; only the per-layer times matter.
mkdir -p gpurun_out
L="conv2_3x3_reduce,conv2_3x3,inception_3a_3x3,inception_3a_double_3x3_2,inception_4a_1x1,inception_4a_double_3x3_2,inception_4e_double_3x3_1"
for cfg in "SSNB_V2=0" "SSNB_V2=1"; do
  for ab in 0 1 4 8 12; do
    env $cfg SSNB_ABLATE=$ab SSNB_LAYERS=$L timeout 200 python tools/layer_times.py 288 > gpurun_out/abl_${cfg}_$ab.txt 2>&1
    echo "$cfg ablate=$ab: $(grep -v TOTAL gpurun_out/abl_${cfg}_$ab.txt | awk '{printf "%s ", $(NF-4)}')"
  done
done
for ab in 28 44 60; do
  env SSNB_V2=0 SSNB_ABLATE=$ab SSNB_LAYERS=$L timeout 200 python tools/layer_times.py 288 > gpurun_out/abl_gen1_$ab.txt 2>&1
  echo "SSNB_V2=0 ablate=$ab: $(grep -v TOTAL gpurun_out/abl_gen1_$ab.txt | awk '{printf "%s ", $(NF-4)}')"
done
