#!/bin/bash
# pair_tz kernel (four time steps per GEMM row, C = 32 stage): device unit tests (watchdog flavour), perf shapes next to
# the pair kernel it replaces, cycle trace of two shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/tz_units.log
: > $L
for i in 0 1 2 3 4 5 6 7 8 9 10 11; do timeout 120 build/test_pair_tz_wd $i 1 >> $L 2>&1; echo "rc=$?" >> $L; done
grep -E "PASS|FAIL|rc=[^0]|WATCHDOG|error|mismatch" $L | head -60
P=gpurun_out/tz_perf.log
: > $P
for i in 12 13 14 15 16 17 18 19; do timeout 120 build/test_pair_tz $i 5 >> $P 2>&1; done
for i in 14 16 17 18 19 20 15; do timeout 120 build/test_pair_tc $i 5 >> $P 2>&1; done
for i in 16 17 12; do timeout 120 build/test_pair_tz_trace $i 2 2>&1 | grep -E "^\[tz|trace" | awk '!seen[$0]++' >> $P; done
for m in 1 2 4 7; do timeout 120 build/test_pair_tz 16 5 $m >> $P 2>&1; done
grep -E "^\[.*perf|time|trace|dbg" $P | cut -c1-300
