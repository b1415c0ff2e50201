#!/bin/bash
# Runs on the GPU box: per-launch device times of one bench run (ncu, serialised) -> gpurun_out/launches_$1.csv
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_$1.csv $CMD > gpurun_out/ncu_$1.log 2>&1
tail -n 2 gpurun_out/ncu_$1.log | cut -c1-300
python bench.py --steps 50 --warmup 5 --no-cpu-baseline | cut -c1-200
