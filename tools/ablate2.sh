#!/bin/bash
mkdir -p gpurun_out
L="conv2_3x3_reduce,conv2_3x3,inception_3a_3x3,inception_3a_double_3x3_2,inception_4a_1x1,inception_4a_double_3x3_2,inception_4e_double_3x3_1"
for ab in 12 28 44 60; do
  env SSNB_HALO=0 SSNB_ABLATE=$ab SSNB_LAYERS=$L timeout 200 python tools/layer_times.py 288 > gpurun_out/abl2_$ab.txt 2>&1
  echo "classic ablate=$ab: $(grep -v TOTAL gpurun_out/abl2_$ab.txt | awk '{printf "%s ", $(NF-4)}')"
done
