#!/bin/bash
# `ncu --set full` of selected launches of one training step: the engine brackets the listed ops with
# cudaProfilerStart/Stop (SSNB_PROFILE_FWD_OPS / SSNB_PROFILE_BWD_OPS).  Summarise with tools/ncu_summary.py.
mkdir -p gpurun_out
SSNB_PROFILE_FWD_OPS=${FWD_OPS:-conv2_3x3,inception_3b_pool,pool1_3x3_s2,inception_4a_1x1} \
SSNB_PROFILE_BWD_OPS=${BWD_OPS:-inception_3b_1x1,conv2_3x3} \
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/ncu_targets -f python tools/ncu_step.py 1 ${PREC:-exact_tc} > gpurun_out/ncu_targets.log 2>&1
tail -1 gpurun_out/ncu_targets.log
