#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -k "exact_tc" > gpurun_out/tc_pytest.txt 2>&1; grep -E "passed|failed|rel-L2|worst|aggregate|fused_step" gpurun_out/tc_pytest.txt | cut -c1-400
timeout 600 python bench.py --precision exact_tc --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/tc_bench.json 2> gpurun_out/tc_bench.err; cut -c1-400 gpurun_out/tc_bench.json; tail -3 gpurun_out/tc_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tc_launches.csv python tools/ncu_step.py 3 exact_tc > gpurun_out/tc_ncu_step.log 2>&1
python tools/launch_summary.py gpurun_out/tc_launches.csv | head -40
