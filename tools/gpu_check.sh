#!/bin/bash
# One GPU visit (run under gpurun): tensor-core launches vs the SIMT kernels (FAST and EXACT_TC), the GPU test suite, a bench
# line, and the per-launch time list of one training step.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 400 python tools/umma_diag.py 160 > gpurun_out/diag.txt 2>&1; echo "diag fast: $(grep -c '^BAD' gpurun_out/diag.txt) BAD; $(tail -1 gpurun_out/diag.txt | cut -c1-100)"
timeout 400 python tools/umma_diag.py 18 tc > gpurun_out/diag_tc.txt 2>&1; echo "diag tc: $(grep -c '^BAD' gpurun_out/diag_tc.txt) BAD; $(tail -1 gpurun_out/diag_tc.txt | cut -c1-100)"
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.txt 2>&1; grep -E "passed|failed|rel-L2|aggregate|^FAILED" gpurun_out/pytest_gpu.txt | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python tools/ncu_step.py 3 exact_tc > gpurun_out/ncu_step_tc.log 2>&1
python tools/launch_summary.py gpurun_out/launches_tc.csv
fi
