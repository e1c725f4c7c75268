#!/bin/bash
# after a kernel change (one GPU, ~3 min): every tensor-core launch vs the SIMT kernels in both modes, the bench line with the
# per-kernel times of the roofline leg, the tensor-core parity tests
mkdir -p gpurun_out
timeout 300 python tools/umma_diag.py 18 tc > gpurun_out/sp_diag_tc.txt 2>&1; echo "diag tc: $(grep -c '^BAD' gpurun_out/sp_diag_tc.txt) BAD; $(tail -1 gpurun_out/sp_diag_tc.txt | cut -c1-80)"
timeout 300 python tools/umma_diag.py 160 > gpurun_out/sp_diag_fast.txt 2>&1; echo "diag fast: $(grep -c '^BAD' gpurun_out/sp_diag_fast.txt) BAD; $(tail -1 gpurun_out/sp_diag_fast.txt | cut -c1-80)"
timeout 500 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/sp.json 2> gpurun_out/sp.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/sp.json'))
    print("exact_tc %.2f ms/step (%.1f prop/s) wgrad %.3f ms e2e %.1f | fast %.2f ms (%.1f) | losses %s" % (d["ms_per_step"], d["value"], d["roofline"]["wgrad"]["ms_per_step"], d["e2e"]["value"], d["modes"]["fast"]["ms_per_step"], d["modes"]["fast"]["value"], [round(x,6) for x in d["losses"]]))
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/sp.err').read()[-800:])
PY
timeout 900 python -m pytest tests -m gpu -q -k "exact_tc or fast or bucketed or fused_step or flow" > gpurun_out/sp_tests.txt 2>&1; tail -3 gpurun_out/sp_tests.txt | cut -c1-200
