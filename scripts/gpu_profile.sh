#!/bin/bash
# Runs on the GPU box: launch list of one bench run + ncu --set full of two pair-kernel launches (stage 1 and stage 3, k = 11).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches5.csv $CMD > gpurun_out/ncu5.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pair_tc -s 6 -c 1 -f -o gpurun_out/prof_pair_s1k11 $CMD > gpurun_out/ncu6.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pair_tc -s 24 -c 1 -f -o gpurun_out/prof_pair_s3k11 $CMD > gpurun_out/ncu7.log 2>&1
tail -2 gpurun_out/ncu5.log gpurun_out/ncu6.log gpurun_out/ncu7.log
