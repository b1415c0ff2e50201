#!/bin/bash
# round 2: new mel kernel (shuffle exchanges, frames-on-lanes filterbank): mel tests + timing, then the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.jsonl
timeout 600 python -m pytest tests/test_gpu_mel.py -q -x > gpurun_out/r2d_pytest_mel.log 2>&1; echo "mel pytest rc=$?"; tail -5 gpurun_out/r2d_pytest_mel.log
timeout 300 python scripts/time_mel.py 1024 220500 20 > gpurun_out/r2d_time_mel.log 2>&1; cat gpurun_out/r2d_time_mel.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
