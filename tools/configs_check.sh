#!/bin/bash
# every BASELINE config through bench.py on one GPU (short runs): configs[2] Flow, configs[3] K=200 / 64 proposals, configs[4] inference
mkdir -p gpurun_out
run() { local name=$1; shift
  timeout 1200 python bench.py "$@" > gpurun_out/cfg_$name.json 2> gpurun_out/cfg_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/cfg_$name.json'))
    print("$name: %.1f %s, %.2f ms/step, e2e %.1f, frames/s %s, cpu %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("frames_per_s"), (d.get("cpu_baseline") or {}).get("value")))
    if d.get("modes"): print("   modes:", {k: round(v["value"],1) for k,v in d["modes"].items()})
    if d.get("stpp"): print("   stpp:", json.dumps(d["stpp"])[:700])
except Exception as e:
    print("$name FAILED", e); print(open('gpurun_out/cfg_$name.err').read()[-1500:])
PY
}
run flow --modality Flow --steps 10 --warmup 3
run k200 --classes 200 --videos-per-gpu 8 --steps 10 --warmup 3
run infer --mode infer --steps 5 --warmup 3
run infer_fast --mode infer --precision fast --steps 5 --warmup 3 --no-cpu-baseline
run rgb --steps 10 --warmup 3 --no-cpu-baseline
