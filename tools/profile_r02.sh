#!/bin/bash
# round-2 profiles (run under gpurun, 1 GPU): launch lists of one fused training step in both tensor-core modes, `ncu --set
# full` of selected launches of the EXACT_TC step and of the STPP kernels.  Summaries are written by tools/launch_summary.py /
# tools/ncu_summary.py here; copy them into profiles/.
mkdir -p gpurun_out
for prec in exact_tc fast; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$prec.csv python tools/ncu_step.py 3 $prec > gpurun_out/r02_ncu_step_$prec.log 2>&1
  python tools/launch_summary.py gpurun_out/r02_launches_$prec.csv > gpurun_out/r02_launches_$prec.txt; head -12 gpurun_out/r02_launches_$prec.txt
done
PREC=exact_tc FWD_OPS=conv2_3x3,inception_3b_pool,pool1_3x3_s2,inception_4a_1x1,inception_3a_double_3x3_2 BWD_OPS=inception_3b_1x1,conv2_3x3 bash tools/ncu_targets.sh
python tools/ncu_summary.py gpurun_out/ncu_targets.ncu-rep > gpurun_out/r02_ncu_full_exact_tc.txt 2>&1; grep -c "== launch" gpurun_out/r02_ncu_full_exact_tc.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:stpp -c 6 -o gpurun_out/ncu_stpp -f python tools/ncu_stpp.py exact_tc > gpurun_out/ncu_stpp.log 2>&1
python tools/ncu_summary.py gpurun_out/ncu_stpp.ncu-rep > gpurun_out/r02_ncu_full_stpp.txt 2>&1; grep -c "== launch" gpurun_out/r02_ncu_full_stpp.txt
ls -la gpurun_out/*.ncu-rep
