#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mel.py tests/test_gpu_full_size.py -q -x -k "mel" > gpurun_out/melq_pytest.log 2>&1; echo "mel pytest rc=$?"; tail -4 gpurun_out/melq_pytest.log
timeout 300 python scripts/time_mel.py 1024 220500 20 2>&1 | tee gpurun_out/melq_time.log
grep -E "crop|golden|oracle" gpurun_out/parity_margins.jsonl | tail -12
