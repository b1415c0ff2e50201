#!/bin/bash
# Runs on the GPU box: ncu --set full of the first stage-1 pair (C = 128, k = 3) with the direct and the staged (TMA store) epilogue.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
E2E_PAIR_STAGED=0 ncu --set full --clock-control none --import-source on -k regex:pair_tc -s 0 -c 1 -f -o gpurun_out/prof_pair_s1k3_direct $CMD > gpurun_out/ncu_s0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_tc -s 0 -c 1 -f -o gpurun_out/prof_pair_s1k3_staged $CMD > gpurun_out/ncu_s1.log 2>&1
tail -n 1 gpurun_out/ncu_s0.log gpurun_out/ncu_s1.log
