#!/bin/bash
mkdir -p gpurun_out
SSNB_PAIR=1 timeout 300 python tools/umma_diag.py 18 > gpurun_out/diag_pair.txt 2>&1; echo "diag pair: $(grep -c '^BAD' gpurun_out/diag_pair.txt) BAD; $(tail -1 gpurun_out/diag_pair.txt)"
SSNB_HALO=0 timeout 300 python tools/layer_times.py 288 > gpurun_out/lt_classic.txt 2>&1; echo "classic: $(tail -1 gpurun_out/lt_classic.txt)"
timeout 300 python tools/layer_times.py 288 > gpurun_out/lt_halo.txt 2>&1; echo "halo: $(tail -1 gpurun_out/lt_halo.txt)"
if grep -q "umma_diag: 0 mismatching" gpurun_out/diag_pair.txt; then
  SSNB_PAIR=1 timeout 300 python tools/layer_times.py 288 > gpurun_out/lt_pair.txt 2>&1; echo "pair: $(tail -1 gpurun_out/lt_pair.txt)"
  SSNB_PAIR=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; cut -c1-330 gpurun_out/bench_pair.json
fi
NCU="ncu --set full --import-source on --clock-control none -k regex:umma_conv"
SSNB_HALO=0 timeout 400 $NCU -o gpurun_out/ncu_classic -f python tools/ncu_layers.py > gpurun_out/ncu_classic.log 2>&1; tail -2 gpurun_out/ncu_classic.log
SSNB_HALO=0 SSNB_EPI=1 timeout 400 $NCU -o gpurun_out/ncu_classic_epi1 -f python tools/ncu_layers.py > gpurun_out/ncu_classic_epi1.log 2>&1; tail -1 gpurun_out/ncu_classic_epi1.log
timeout 400 $NCU -o gpurun_out/ncu_halo -f python tools/ncu_layers.py > gpurun_out/ncu_halo.log 2>&1; tail -1 gpurun_out/ncu_halo.log
if grep -q "umma_diag: 0 mismatching" gpurun_out/diag_pair.txt; then
  SSNB_PAIR=1 timeout 400 $NCU -o gpurun_out/ncu_pair -f python tools/ncu_layers.py > gpurun_out/ncu_pair.log 2>&1; tail -1 gpurun_out/ncu_pair.log
fi
ls -la gpurun_out/*.ncu-rep
