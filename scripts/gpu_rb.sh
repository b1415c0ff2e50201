#!/bin/bash
# fused whole-resblock kernel: device unit tests (watchdog flavour), perf shapes, then the vocoder parity tests + a quick bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/rb_units.log
: > $L
for i in 0 1 2 3 4 5 6 7 8; do timeout 120 build/test_rb_tc_wd $i 1 >> $L 2>&1; echo "rc=$?" >> $L; done
grep -E "PASS|FAIL|rc=[^0]|WATCHDOG|error" $L | head -40
for i in 9 10 11 12 13; do timeout 120 build/test_rb_tc $i 5 >> gpurun_out/rb_perf.log 2>&1; done
for i in 9 12 14; do timeout 120 build/test_pair_tc $i 5 >> gpurun_out/rb_perf.log 2>&1; done
grep -E "^\[|time" gpurun_out/rb_perf.log
timeout 900 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_full_size.py -x -q > gpurun_out/rb_pytest.log 2>&1; tail -5 gpurun_out/rb_pytest.log
timeout 300 python bench.py --quick --no-side > gpurun_out/rb_bench.json 2> gpurun_out/rb_bench.err; head -c 600 gpurun_out/rb_bench.json
E2E_RB_FUSION=0 timeout 300 python bench.py --quick --no-side > gpurun_out/rb_bench_off.json 2>> gpurun_out/rb_bench.err; head -c 600 gpurun_out/rb_bench_off.json
