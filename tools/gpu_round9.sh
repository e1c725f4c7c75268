#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/umma_diag.py 160 > gpurun_out/diag.txt 2>&1; echo "diag: $(grep -c '^BAD' gpurun_out/diag.txt) BAD; $(tail -1 gpurun_out/diag.txt | cut -c1-100)"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-230 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py 3 > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
SSNB_PROFILE_FWD_OPS=conv2_3x3,inception_3b_pool,pool1_3x3_s2,inception_4a_1x1 SSNB_PROFILE_BWD_OPS=inception_3b_1x1,conv2_3x3 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/ncu_targets2 -f python tools/ncu_step.py 1 > gpurun_out/ncu_targets2.log 2>&1; tail -1 gpurun_out/ncu_targets2.log; ls -la gpurun_out/ncu_targets2.ncu-rep
