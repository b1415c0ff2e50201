#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/time_b1.py 1 50 | tee gpurun_out/b1_time.log
python scripts/time_mel.py 1024 220500 20 | tee -a gpurun_out/b1_time.log
python scripts/time_b1.py 1 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_b1.csv python scripts/time_b1.py 1 3 > gpurun_out/ncu_b1.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_b1.csv 4 | tee gpurun_out/b1_launches.txt | tail -60
