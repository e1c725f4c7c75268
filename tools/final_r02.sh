#!/bin/bash
# end-of-round check (1 GPU): launch lists of one fused training step in both tensor-core modes (-> profiles/), the whole
# `-m gpu` suite, the default bench line
mkdir -p gpurun_out
for prec in exact_tc fast; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$prec.csv python tools/ncu_step.py 3 $prec > gpurun_out/r02_ncu_step_$prec.log 2>&1
  python tools/launch_summary.py gpurun_out/r02_launches_$prec.csv > gpurun_out/r02_launches_$prec.txt; head -4 gpurun_out/r02_launches_$prec.txt
done
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; grep -E "passed|failed|^FAILED|Error" gpurun_out/pytest_gpu.txt | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print("default: %.1f prop/s %.2f ms; e2e %.1f; fast %.1f (%.2f ms); cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["modes"]["fast"]["value"], d["modes"]["fast"]["ms_per_step"], d["cpu_baseline"]))
r=d["roofline"]
print("roofline frac %.3f tensor_pipe_frac %.3f; fwd %.1f dgrad %.1f wgrad %.1f TF/s; stpp large fwd %.3f bwd %.3f fused %.3f" % (r["frac"], r["tensor_pipe_frac"], r["forward"]["achieved"], r["dgrad"]["achieved"], r["wgrad"]["achieved"], d["stpp"]["large"]["fwd"]["frac_of_hbm_peak"], d["stpp"]["large"]["bwd"]["frac_of_hbm_peak"], d["stpp"]["fused_gpool_stpp"]["frac_of_hbm_peak"]))
PY
tail -2 gpurun_out/bench.err
