#!/bin/bash
# GPU probe: which realisation of the shifted halo views the UMMA descriptors accept (SSNB_HALO_MODE 0..4), checked
# layer by layer against the SIMT kernels; then per-layer timings of the first exact mode and of the classic path.
mkdir -p gpurun_out
for m in 2 3 0 1 4; do
  SSNB_HALO_MODE=$m timeout 300 python tools/umma_diag.py 18 > gpurun_out/halo_diag_mode$m.txt 2>&1
  echo "mode $m: $(grep -c '^BAD' gpurun_out/halo_diag_mode$m.txt) BAD lines; $(tail -1 gpurun_out/halo_diag_mode$m.txt)"
done
best=""
for m in 2 0 3 1 4; do
  if grep -q "umma_diag: 0 mismatching" gpurun_out/halo_diag_mode$m.txt; then best=$m; break; fi
done
echo "best mode: '$best'"
SSNB_HALO=0 timeout 300 python tools/layer_times.py 288 > gpurun_out/layer_times_classic.txt 2>&1; tail -1 gpurun_out/layer_times_classic.txt
if [ -n "$best" ]; then
  for m in 2 0 4; do
    if grep -q "umma_diag: 0 mismatching" gpurun_out/halo_diag_mode$m.txt; then
      SSNB_HALO_MODE=$m timeout 300 python tools/layer_times.py 288 > gpurun_out/layer_times_halo$m.txt 2>&1; echo "halo mode $m: $(tail -1 gpurun_out/layer_times_halo$m.txt)"
    fi
  done
  SSNB_HALO_MODE=$best timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_halo.json 2> gpurun_out/bench_halo.err; cat gpurun_out/bench_halo.json | cut -c1-400
fi
