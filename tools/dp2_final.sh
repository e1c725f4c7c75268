#!/bin/bash
# 2-GPU check of the final code: the NCCL equivalence test and the bench line in both tensor-core modes
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dp_nccl.py -m gpu -q -s > gpurun_out/dp_nccl.txt 2>&1; grep -E "passed|failed|Error|assert" gpurun_out/dp_nccl.txt | cut -c1-300
for prec in exact_tc fast; do
SSNB_NCCL_TIMEOUT_S=90 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --precision $prec --no-second-mode --no-cpu-baseline > gpurun_out/bench_n2_$prec.json 2> gpurun_out/bench_n2_$prec.err
python - <<PY
import json
try:
    txt=open('gpurun_out/bench_n2_$prec.json').read()
    d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("$prec N=2: %.1f prop/s %.3f ms/step e2e %.1f graph=%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("cuda_graph")))
except Exception as e:
    print("$prec N=2 failed", e); print(open('gpurun_out/bench_n2_$prec.err').read()[-1500:])
PY
done
