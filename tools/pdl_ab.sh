#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/umma_diag.py 18 tc > gpurun_out/pdl_diag_tc.txt 2>&1; echo "diag tc (PDL on): $(grep -c '^BAD' gpurun_out/pdl_diag_tc.txt) BAD; $(tail -1 gpurun_out/pdl_diag_tc.txt | cut -c1-80)"
timeout 300 python tools/umma_diag.py 160 > gpurun_out/pdl_diag_fast.txt 2>&1; echo "diag fast (PDL on): $(grep -c '^BAD' gpurun_out/pdl_diag_fast.txt) BAD; $(tail -1 gpurun_out/pdl_diag_fast.txt | cut -c1-80)"
for q in 1 0; do
  SSNB_PDL=$q timeout 500 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/pdl_${q}.json 2> gpurun_out/pdl_${q}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/pdl_$q.json'))
    print("PDL=$q: exact_tc %.2f ms/step (%.1f prop/s) graph=%s e2e %.1f | fast %.2f ms (%.1f) | losses %s" % (d["ms_per_step"], d["value"], d["config"]["cuda_graph"], d["e2e"]["value"], d["modes"]["fast"]["ms_per_step"], d["modes"]["fast"]["value"], [round(x,6) for x in d["losses"]]))
except Exception as e:
    print("PDL=$q failed", e); print(open('gpurun_out/pdl_$q.err').read()[-800:])
PY
done
timeout 900 python -m pytest tests -m gpu -q -k "exact_tc or fast or bucketed" > gpurun_out/pdl_tests.txt 2>&1; tail -3 gpurun_out/pdl_tests.txt | cut -c1-200
