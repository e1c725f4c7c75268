#!/bin/bash
# A/B of the weight-gradient x staging (classic per-tap boxes vs halo boxes) in EXACT_TC, where every tile is staged three times
mkdir -p gpurun_out
for h in 3 1 2; do
  SSNB_WGRAD_HALO=$h timeout 300 python tools/umma_diag.py 18 tc > gpurun_out/wh_${h}_diag.txt 2>&1
  SSNB_WGRAD_HALO=$h timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-second-mode > gpurun_out/wh_${h}.json 2> gpurun_out/wh_${h}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/wh_$h.json'))
    print("WGRAD_HALO=$h: BAD=%s  %.2f ms/step  wgrad %.3f ms" % (open('gpurun_out/wh_${h}_diag.txt').read().count("\nBAD"), d["ms_per_step"], d["roofline"]["wgrad"]["ms_per_step"]))
except Exception as e:
    print("WGRAD_HALO=$h failed", e)
PY
done
