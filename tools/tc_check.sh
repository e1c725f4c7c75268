#!/bin/bash
# EXACT_TC bring-up (run under gpurun): every split-operand launch vs the fp32 SIMT kernels, then the parity tests.
mkdir -p gpurun_out
timeout 600 python tools/umma_diag.py 18 tc > gpurun_out/tc_diag18.txt 2>&1; echo "diag18: $(grep -c '^BAD' gpurun_out/tc_diag18.txt) BAD; $(tail -1 gpurun_out/tc_diag18.txt | cut -c1-120)"
timeout 600 python tools/umma_diag.py 160 tc > gpurun_out/tc_diag160.txt 2>&1; echo "diag160: $(grep -c '^BAD' gpurun_out/tc_diag160.txt) BAD; $(tail -1 gpurun_out/tc_diag160.txt | cut -c1-120)"
timeout 1200 python -m pytest tests -m gpu -q -s -k "exact_tc" > gpurun_out/tc_pytest.txt 2>&1; tail -15 gpurun_out/tc_pytest.txt
