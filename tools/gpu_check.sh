#!/bin/bash
# One GPU visit (run under gpurun): every tcgen05 launch against the SIMT kernels at a multi-tile shape, the GPU test
# suite, a bench line, and the per-launch time list of one training step.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 400 python tools/umma_diag.py 160 > gpurun_out/diag.txt 2>&1; echo "diag: $(grep -c '^BAD' gpurun_out/diag.txt) BAD; $(tail -1 gpurun_out/diag.txt | cut -c1-100)"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-230 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py 3 > gpurun_out/ncu_step.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv
