#!/bin/bash
# full GPU suite + every BASELINE config through bench.py (1 GPU)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.txt 2>&1; grep -E "passed|failed|^FAILED|Error|bn partial|fast tcgen05" gpurun_out/pytest_gpu.txt | cut -c1-300
bash tools/configs_check.sh
