#!/bin/bash
# Runs on the GPU box: CTA-pair form of the unfused convolution kernel - correctness (watchdog flavour) then timing.
fail=0
for i in 5 6 25 22 24; do E2E_CONV_CG=2 timeout 60 ./build/test_conv_tc_wd $i 3 | grep -E "cfg|plan|err|PASS|FAIL|WATCHDOG|error" || fail=1; done
echo "---- timing"
for i in 16 22 23 24; do for cg in 1 2; do E2E_CONV_CG=$cg timeout 60 ./build/test_conv_tc $i 20 | grep -E "plan|time|FAIL" | cut -c1-120; done; done
