#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -4 gpurun_out/smoke.txt
timeout 600 python -m pytest tests/test_dp_nccl.py -m gpu -q -s > gpurun_out/dp_nccl.txt 2>&1; grep -E "passed|failed|Error|assert" gpurun_out/dp_nccl.txt | cut -c1-300
for prec in exact_tc fast; do
SSNB_NCCL_TIMEOUT_S=90 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --precision $prec --no-second-mode > gpurun_out/bench_n2_$prec.json 2> gpurun_out/bench_n2_$prec.err
python - <<PY
import json
try:
    txt=open('gpurun_out/bench_n2_$prec.json').read()
    d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("$prec N=2: %.1f prop/s %.3f ms/step e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$prec N=2 failed", e); print(open('gpurun_out/bench_n2_$prec.err').read()[-1500:])
PY
done
SSNB_NCCL_TIMEOUT_S=90 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; cut -c1-200 gpurun_out/bench_n2_ref.json
