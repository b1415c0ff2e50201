#!/bin/bash
# Standalone device tests (run on the GPU box by scripts/gpu_*.sh): production flavour, watchdog flavour
# (-DE2E_WATCHDOG: every mbarrier spin is clock-bounded and traps with a call-site code) and traced flavour.
set -e
cd "$(dirname "$0")/.."
mkdir -p build
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17"
pids=()
for t in conv_tc pair_tc rb_tc pair_tz; do
  nvcc $F -o build/test_$t tests/cuda/test_$t.cu & pids+=($!)
  nvcc $F -DE2E_WATCHDOG -o build/test_${t}_wd tests/cuda/test_$t.cu & pids+=($!)
done
nvcc $F -DE2E_TRACE -o build/test_pair_tc_trace tests/cuda/test_pair_tc.cu & pids+=($!)
nvcc $F -DE2E_TRACE2 -o build/test_pair_tc_trace2 tests/cuda/test_pair_tc.cu & pids+=($!)
nvcc $F -DE2E_TZTRACE -o build/test_pair_tz_trace tests/cuda/test_pair_tz.cu & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
ls -la build
