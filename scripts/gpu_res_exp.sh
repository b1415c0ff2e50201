#!/bin/bash
# what do the residual's global loads cost in pair_tc (upper bound of an x-in-TMEM variant)?  + compute-sanitizer on pair_tz
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/res_exp.log
: > $P
for i in 9 10 11 12 13; do for m in 0 1; do timeout 120 build/test_pair_tc $i 5 $m >> $P 2>&1; done; done
grep -E "^\[pair .*perf|time|dbg" $P
timeout 300 compute-sanitizer --tool memcheck build/test_pair_tz 0 1 > gpurun_out/sanitizer_tz.log 2>&1; tail -5 gpurun_out/sanitizer_tz.log
timeout 300 compute-sanitizer --tool memcheck build/test_pair_tz 9 1 >> gpurun_out/sanitizer_tz.log 2>&1; tail -4 gpurun_out/sanitizer_tz.log
