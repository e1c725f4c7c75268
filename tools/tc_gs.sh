#!/bin/bash
mkdir -p gpurun_out
for gs in 1024 65536 4194304; do
  SSNB_TEST_GS=$gs timeout 900 python -m pytest tests -m gpu -q -s -k "test_ssn_train_exact_vs_oracle and exact_tc" > gpurun_out/tc_gs_$gs.txt 2>&1
  echo "gs=$gs: $(grep -E 'aggregate|overflow|passed|failed' gpurun_out/tc_gs_$gs.txt | cut -c1-200 | tr '\n' '|')"
done
