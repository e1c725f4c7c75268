#!/bin/bash
# the other BASELINE configs on the final code (GPU legs only; the CPU reference numbers are those of tools/configs_check.sh)
mkdir -p gpurun_out
run() { local name=$1; shift
  timeout 400 python bench.py "$@" --no-cpu-baseline > gpurun_out/cfg_$name.json 2> gpurun_out/cfg_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/cfg_$name.json'))
    print("$name: %.1f %s, %.2f ms/step, e2e %.1f, frames/s %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("frames_per_s")))
    if d.get("modes"): print("   modes:", {k: (round(v["value"],1), round(v["ms_per_step"],2)) for k,v in d["modes"].items()})
except Exception as e:
    print("$name FAILED", e); print(open('gpurun_out/cfg_$name.err').read()[-1500:])
PY
}
run flow --modality Flow --steps 10 --warmup 3
run k200 --classes 200 --videos-per-gpu 8 --steps 10 --warmup 3
run infer --mode infer --steps 5 --warmup 3
run infer_fast --mode infer --precision fast --steps 5 --warmup 3
