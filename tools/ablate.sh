#!/bin/bash
# timing experiment: which role bounds the conv kernel (SSNB_ABLATE bits: 1 no stores, 2 no bias loads, 4 empty epilogue, 8 no MMAs)
mkdir -p gpurun_out
L="conv2_3x3_reduce,conv2_3x3,inception_3a_3x3,inception_3a_double_3x3_2,inception_4a_1x1,inception_4a_double_3x3_2,inception_4e_double_3x3_1"
for cfg in "SSNB_HALO=0" "SSNB_HALO=1" "SSNB_PAIR=1"; do
  for ab in 0 2 1 3 4 8 12; do
    env $cfg SSNB_ABLATE=$ab SSNB_LAYERS=$L timeout 200 python tools/layer_times.py 288 > gpurun_out/abl_${cfg}_$ab.txt 2>&1
    echo "$cfg ablate=$ab: $(grep -v TOTAL gpurun_out/abl_${cfg}_$ab.txt | awk '{printf "%s ", $(NF-4)}')"
  done
done
