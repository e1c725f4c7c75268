#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/umma_diag.py 160 > gpurun_out/diag_v2.txt 2>&1; echo "diag v2: $(grep -c '^BAD' gpurun_out/diag_v2.txt) BAD; $(tail -1 gpurun_out/diag_v2.txt)"
SSNB_PAIR=1 timeout 300 python tools/umma_diag.py 160 > gpurun_out/diag_v2pair.txt 2>&1; echo "diag v2 pair: $(grep -c '^BAD' gpurun_out/diag_v2pair.txt) BAD; $(tail -1 gpurun_out/diag_v2pair.txt)"
timeout 300 python tools/layer_times.py 288 > gpurun_out/lt_v2.txt 2>&1; echo "v2: $(tail -1 gpurun_out/lt_v2.txt)"
SSNB_PAIR=1 timeout 300 python tools/layer_times.py 288 > gpurun_out/lt_v2pair.txt 2>&1; echo "v2 pair: $(tail -1 gpurun_out/lt_v2pair.txt)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; cut -c1-330 gpurun_out/bench_v2.json
SSNB_PAIR=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2pair.json 2> gpurun_out/bench_v2pair.err; cut -c1-330 gpurun_out/bench_v2pair.json
