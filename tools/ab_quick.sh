#!/bin/bash
# Lean A/B of the default-off kernel variants (run under gpurun): every tcgen05 launch vs the SIMT kernels, then a bench line.
mkdir -p gpurun_out
run() {
  local name=$1; shift
  env "$@" timeout 300 python tools/umma_diag.py 160 > gpurun_out/ab_${name}_diag.txt 2>&1
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${name}_bench.json 2> gpurun_out/ab_${name}_bench.err
  echo "$name: BAD=$(grep -c '^BAD' gpurun_out/ab_${name}_diag.txt) | $(tail -1 gpurun_out/ab_${name}_diag.txt | cut -c1-80) | $(python -c "import json,sys; d=json.load(open('gpurun_out/ab_${name}_bench.json')); print('%.3f ms/step, e2e %.0f' % (d['ms_per_step'], d['e2e']['value']))" 2>&1 | tail -1)"
}
run default SSNB_NOP=1
run epi_tma SSNB_EPI_TMA=1
run epi_deep SSNB_EPI_DEEP=1
run avgpool_pair SSNB_AVGPOOL=pair
run lib_graph SSNB_GRAPH=1
