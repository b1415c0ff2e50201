#!/bin/bash
# Runs on the GPU box: standalone kernel tests (one process per configuration), pytest -m gpu, bench.
mkdir -p gpurun_out
fail=0
n=$(./build/test_pair_tc list); for i in $(seq 0 $((n-1))); do timeout 120 ./build/test_pair_tc $i 20 || fail=1; done > gpurun_out/pair_all.log 2>&1
n=$(./build/test_conv_tc list); for i in $(seq 0 $((n-1))); do timeout 120 ./build/test_conv_tc $i 20 || fail=1; done > gpurun_out/conv_all.log 2>&1
grep -E "PASS|FAIL|time|WATCHDOG|error" gpurun_out/pair_all.log gpurun_out/conv_all.log | grep -E "FAIL|WATCHDOG|error|perf" -A1 | head -60
echo "unit fail=$fail"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -5 gpurun_out/pytest_gpu_all.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-400
