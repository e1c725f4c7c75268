#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/umma_diag.py 160 > gpurun_out/diag.txt 2>&1; echo "diag: $(grep -c '^BAD' gpurun_out/diag.txt) BAD; $(tail -1 gpurun_out/diag.txt)"
SSNB_EPI=1 timeout 400 python tools/umma_diag.py 160 > gpurun_out/diag_epi1.txt 2>&1; echo "diag epi1: $(grep -c '^BAD' gpurun_out/diag_epi1.txt) BAD; $(tail -1 gpurun_out/diag_epi1.txt)"
SSNB_EPI=0 timeout 300 python tools/layer_times.py 288 > gpurun_out/lt_epi0.txt 2>&1; echo "lt epi0: $(tail -1 gpurun_out/lt_epi0.txt)"
SSNB_EPI=1 timeout 300 python tools/layer_times.py 288 > gpurun_out/lt_epi1.txt 2>&1; echo "lt epi1: $(tail -1 gpurun_out/lt_epi1.txt)"
for e in auto 0 1; do
  if [ $e = auto ]; then unset SSNB_EPI; else export SSNB_EPI=$e; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_epi$e.json 2> gpurun_out/bench_epi$e.err; echo "epi $e: $(cut -c1-200 gpurun_out/bench_epi$e.json)"
done
unset SSNB_EPI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py 3 > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
