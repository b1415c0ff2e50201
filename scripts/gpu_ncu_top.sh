#!/bin/bash
# ncu --set full of the dominant launch (stage-1 k = 11 pair: the 4th pair_tc launch of a forward) and of the C = 64 fused resblock
cd "$(dirname "$0")/.."
TAG=${1:-f}
CMD="python bench.py --steps 1 --warmup 3 --passes 1 --quick --no-side"
$CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pair_tc -s 3 -c 1 -f -o gpurun_out/prof_pair_s1k11_$TAG $CMD > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rb_tc -s 1 -c 1 -f -o gpurun_out/prof_rb_s2k3_$TAG $CMD > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/prof_*_$TAG.ncu-rep
