#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/time_mel.py 1024 220500 3 > gpurun_out/ncu_mel_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mel_kernel -s 3 -c 1 -f -o gpurun_out/prof_mel_r2 python scripts/time_mel.py 1024 220500 3 > gpurun_out/ncu_mel.log 2>&1
tail -3 gpurun_out/ncu_mel.log
