#!/bin/bash
# A/B of the experimental (default-off) variants, run under gpurun: correctness first (tensor-core launches vs the SIMT
# kernels, per-layer parity tests), then a bench line each.  Outputs in gpurun_out/ab_*.
mkdir -p gpurun_out
run() {   # name, env assignments...
  local name=$1; shift
  env "$@" timeout 400 python tools/umma_diag.py 160 > gpurun_out/ab_${name}_diag.txt 2>&1
  env "$@" timeout 600 python -m pytest tests -m gpu -x -q -k "per_layer or fused_vs_unfused or ssn_train" > gpurun_out/ab_${name}_pytest.txt 2>&1
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${name}_bench.json 2> gpurun_out/ab_${name}_bench.err
  echo "$name: $(tail -1 gpurun_out/ab_${name}_diag.txt | cut -c1-60) | $(tail -1 gpurun_out/ab_${name}_pytest.txt) | $(python -c "import json,sys; d=json.load(open('gpurun_out/ab_${name}_bench.json')); print('%.3f ms/step, e2e %.0f' % (d['ms_per_step'], d['e2e']['value']))" 2>/dev/null)"
}
run default SSNB_NOP=1
run epi_deep SSNB_EPI_DEEP=1
run epi_tma SSNB_EPI_TMA=1
run lib_graph SSNB_GRAPH=1
run avgpool_pair SSNB_AVGPOOL=pair
run all SSNB_EPI_TMA=1 SSNB_AVGPOOL=pair
