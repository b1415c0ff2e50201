#!/bin/bash
# Runs on the GPU box: standalone kernel tests (watchdog flavour for the small shapes first, then the production
# flavour for every configuration, one process each), the traced perf shapes, pytest -m gpu, bench.
mkdir -p gpurun_out
fail=0
: > gpurun_out/pair_all.log; : > gpurun_out/conv_all.log
n=$(./build/test_pair_tc list)
for i in $(seq 0 8); do timeout 60 ./build/test_pair_tc_wd $i 3 >> gpurun_out/pair_all.log 2>&1 || fail=1; done
if [ $fail = 0 ]; then for i in $(seq 0 $((n-1))); do timeout 120 ./build/test_pair_tc $i 20 >> gpurun_out/pair_all.log 2>&1 || fail=1; done; fi
n=$(./build/test_conv_tc list)
for i in 0 1 2 3; do timeout 60 ./build/test_conv_tc_wd $i 3 >> gpurun_out/conv_all.log 2>&1 || fail=1; done
if [ $fail = 0 ]; then for i in $(seq 0 $((n-1))); do timeout 120 ./build/test_conv_tc $i 20 >> gpurun_out/conv_all.log 2>&1 || fail=1; done; fi
grep -E "FAIL|WATCHDOG|error" gpurun_out/pair_all.log gpurun_out/conv_all.log | head -20
grep -E "perf" -A3 gpurun_out/pair_all.log | grep -E "perf|time" | paste - - | sed 's/  */ /g' | cut -c1-150
echo "unit fail=$fail"
[ $fail = 0 ] || exit 1
for i in 9 11 12 14 15; do timeout 120 ./build/test_pair_tc_trace $i 20; done > gpurun_out/pair_trace.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -n 3 gpurun_out/pytest_gpu_all.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-330 gpurun_out/bench.json
