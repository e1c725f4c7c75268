#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/umma_diag.py 160 > gpurun_out/diag.txt 2>&1; echo "diag: $(grep -c '^BAD' gpurun_out/diag.txt) BAD; $(tail -1 gpurun_out/diag.txt)"
if ! grep -q "umma_diag: 0 mismatching" gpurun_out/diag.txt; then
  grep "^BAD" gpurun_out/diag.txt | head -20
  export SSNB_WGRAD_HALO=0
  timeout 400 python tools/umma_diag.py 160 > gpurun_out/diag_nohalo.txt 2>&1; echo "diag (wgrad halo off): $(grep -c '^BAD' gpurun_out/diag_nohalo.txt) BAD; $(tail -1 gpurun_out/diag_nohalo.txt)"
fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-230 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py 3 > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log
SSNB_PROFILE_FWD_OPS=conv2_3x3,inception_3b_pool,pool1_3x3_s2 SSNB_PROFILE_BWD_OPS=inception_3b_1x1,conv2_3x3 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/ncu_targets -f python tools/ncu_step.py 1 > gpurun_out/ncu_targets.log 2>&1; tail -2 gpurun_out/ncu_targets.log; ls -la gpurun_out/ncu_targets.ncu-rep
