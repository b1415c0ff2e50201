#!/bin/bash
# rb_tz: where the per-unit time goes (trace flavour) and what the seed / slab stores / global stores cost (dbg masks)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/tz_trace.log
: > $P
for i in 16 17 13 12; do timeout 120 build/test_rb_tz_trace $i 3 >> $P 2>&1; done
for m in 1 2 4 7; do timeout 120 build/test_rb_tz 16 5 $m >> $P 2>&1; done
for m in 4 7; do timeout 120 build/test_rb_tz 17 5 $m >> $P 2>&1; done
grep -E "^\[tz .*perf|time|trace|dbg" $P
