#!/bin/bash
# One GPU visit: layer-by-layer diag, per-layer timings (halo on/off), short bench, GPU test-suite.
mkdir -p gpurun_out
timeout 300 python tools/umma_diag.py 18 > gpurun_out/diag.txt 2>&1; echo "diag: $(grep -c '^BAD' gpurun_out/diag.txt) BAD; $(tail -1 gpurun_out/diag.txt)"
SSNB_HALO=0 timeout 300 python tools/layer_times.py 288 > gpurun_out/layer_times_classic.txt 2>&1; echo "classic: $(tail -1 gpurun_out/layer_times_classic.txt)"
timeout 300 python tools/layer_times.py 288 > gpurun_out/layer_times_halo.txt 2>&1; echo "halo: $(tail -1 gpurun_out/layer_times_halo.txt)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-330 gpurun_out/bench.json
if [ "$1" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -5 gpurun_out/pytest_gpu.txt
fi
