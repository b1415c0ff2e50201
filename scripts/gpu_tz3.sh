#!/bin/bash
# pair_tz in the generator: device units, perf shapes, vocoder parity tests, same-box A/B of the whole forward (E2E_TZ=0 = old plan)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/tz_units.log
: > $L
for i in 0 1 2 3 4 5 6 7 8 9 10 11; do timeout 120 build/test_pair_tz_wd $i 1 >> $L 2>&1; echo "rc=$?" >> $L; done
grep -E "PASS|FAIL|rc=[^0]|WATCHDOG|error|mismatch" $L | head -40
P=gpurun_out/tz_perf.log
: > $P
for i in 12 13 14 15 16 17 18 19; do timeout 120 build/test_pair_tz $i 5 >> $P 2>&1; done
for i in 16 17; do timeout 120 build/test_pair_tz_trace $i 2 2>&1 | grep -E "^\[tz|trace" | awk '!seen[$0]++' | head -3 >> $P; done
grep -E "^\[.*perf|time|trace|dbg" $P | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/tz_pytest.log 2>&1; tail -5 gpurun_out/tz_pytest.log
NOBENCH= bash scripts/gpu_quick_ab.sh "-" "E2E_TZ=0"
