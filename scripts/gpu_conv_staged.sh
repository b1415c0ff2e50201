#!/bin/bash
# Runs on the GPU box: staged (TMA store) epilogue of conv_tc: correctness on the small configurations (watchdog build),
# then the perf configurations with and without it.
mkdir -p gpurun_out
: > gpurun_out/conv_staged.log
fail=0
for i in 1 2 3 4 5 6 7 10 11 12 13 14 25; do E2E_CONV_STAGED=1 timeout 60 ./build/test_conv_tc_wd $i 3 >> gpurun_out/conv_staged.log 2>&1 || { fail=1; echo "cfg $i FAILED"; }; done
grep -E "plan:|PASS|FAIL|WATCHDOG|error" gpurun_out/conv_staged.log | cut -c1-150 | paste - - | cut -c1-220
echo "fail=$fail"
[ $fail = 0 ] || exit 1
for i in 15 16 17 19 20 22 23 24; do
  for st in 0 1; do E2E_CONV_STAGED=$st timeout 120 ./build/test_conv_tc $i 20 2>&1 | grep -E "^\[cfg|plan:|time|FAIL" | tr '\n' ' ' | sed 's/  */ /g' | cut -c1-330; echo; done
done | tee gpurun_out/conv_staged_perf.log
E2E_CONV_STAGED=1 timeout 60 ./build/test_conv_tc_trace 22 20 | tail -n 9
