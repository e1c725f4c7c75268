#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/umma_diag.py 160 > gpurun_out/diag.txt 2>&1; echo "diag: $(grep -c '^BAD' gpurun_out/diag.txt) BAD; $(tail -1 gpurun_out/diag.txt)"
SSNB_HALO_MIN_W=7 timeout 400 python tools/umma_diag.py 160 > gpurun_out/diag_w7.txt 2>&1; echo "diag w7: $(grep -c '^BAD' gpurun_out/diag_w7.txt) BAD; $(tail -1 gpurun_out/diag_w7.txt)"
timeout 300 python tools/layer_times.py 288 > gpurun_out/lt.txt 2>&1; echo "lt: $(tail -1 gpurun_out/lt.txt)"
SSNB_HALO_MIN_W=7 timeout 300 python tools/layer_times.py 288 > gpurun_out/lt_w7.txt 2>&1; echo "lt w7: $(tail -1 gpurun_out/lt_w7.txt)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-330 gpurun_out/bench.json
SSNB_HALO_MIN_W=7 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w7.json 2> gpurun_out/bench_w7.err; cut -c1-330 gpurun_out/bench_w7.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py 3 > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
