#!/bin/bash
# Runs on the GPU box: mel parity tests, config-5 timing, and one ncu --set full capture of the mel kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mel.py -m gpu -x -q 2>&1 | tail -n 5
python scripts/time_mel.py 1024 220500 20 | tee gpurun_out/time_mel.log
python scripts/time_mel.py 64 220500 20 | tee -a gpurun_out/time_mel.log
if [ -n "$NCU" ]; then
ncu --set full --clock-control none --import-source on -k regex:mel_kernel -s 2 -c 1 -f -o gpurun_out/prof_mel python scripts/time_mel.py 1024 220500 2 > gpurun_out/ncu_mel.log 2>&1; tail -n 2 gpurun_out/ncu_mel.log
fi
