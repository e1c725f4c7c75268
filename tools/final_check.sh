#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.txt 2>&1; grep -E "passed|failed|^FAILED|Error" gpurun_out/pytest_gpu.txt | cut -c1-300
timeout 600 python bench.py --precision exact --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err; python -c "
import json; d=json.load(open('gpurun_out/bench_exact.json')); print('exact (SIMT): %.1f prop/s %.1f ms; roofline frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))" || tail -5 gpurun_out/bench_exact.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print("default: %.1f prop/s %.2f ms; e2e %.1f; fast %.1f; cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["modes"]["fast"]["value"], d["cpu_baseline"]))
print("roofline frac %.3f tensor_pipe_frac %.3f; stpp large fwd %.3f bwd %.3f fused %.3f" % (d["roofline"]["frac"], d["roofline"]["tensor_pipe_frac"], d["stpp"]["large"]["fwd"]["frac_of_hbm_peak"], d["stpp"]["large"]["bwd"]["frac_of_hbm_peak"], d["stpp"]["fused_gpool_stpp"]["frac_of_hbm_peak"]))
PY
tail -2 gpurun_out/bench.err
