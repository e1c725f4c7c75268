#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/umma_diag.py 18 tc > gpurun_out/diag_tc.txt 2>&1; echo "diag tc: $(grep -c '^BAD' gpurun_out/diag_tc.txt) BAD; $(tail -1 gpurun_out/diag_tc.txt | cut -c1-100)"
timeout 2400 python -m pytest tests -m gpu -q -s -k "exact_tc or flow_train or fused_sgd or autograd_guards" > gpurun_out/new_tests.txt 2>&1
grep -E "passed|failed|F=288|flow e2e|rel-L2|aggregate|^FAILED|Error|assert " gpurun_out/new_tests.txt | cut -c1-420
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench2.json'))
print("exact_tc %.1f prop/s %.2f ms; e2e %.1f; fast %.1f prop/s %.2f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["modes"]["fast"]["value"], d["modes"]["fast"]["ms_per_step"]))
for k,v in d["roofline"]["top_kernels_ms_per_step"].items(): print("   ", k, v)
PY
tail -3 gpurun_out/bench2.err
