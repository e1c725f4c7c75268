#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/umma_diag.py 160 > gpurun_out/diag.txt 2>&1; echo "diag: $(grep -c '^BAD' gpurun_out/diag.txt) BAD; $(tail -1 gpurun_out/diag.txt | cut -c1-100)"
if ! grep -q "umma_diag: 0 mismatching" gpurun_out/diag.txt; then
  grep -v "^Search\|^CUDA kernel\|^For debug\|^Compile" gpurun_out/diag.txt | grep -v "^ok" | head -20
  export SSNB_WGRAD_HALO=0
fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-230 gpurun_out/bench.json
SSNB_WGRAD_WAVES=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w2.json 2> gpurun_out/bench_w2.err; echo "waves2: $(cut -c1-230 gpurun_out/bench_w2.json)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py 3 > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
