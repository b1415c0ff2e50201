#!/bin/bash
# Runs on the GPU box: the vocoder parity tests under each documented experiment switch - none of them
# may change a result beyond the stated tolerances.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/switch_matrix.log
for s in "E2E_NO_SUM_PREFETCH=1" "E2E_NO_PDL=1" "E2E_NO_GRAPH=1" "E2E_TZ=0" "E2E_RB_FUSION=0" "E2E_NO_TILED_SUMS=1" "E2E_CONV_STAGED=1" "E2E_PAIR_STAGED=1"; do
  r=$(env $s timeout 600 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -n 1)
  echo "$s: $r" | tee -a gpurun_out/switch_matrix.log
done
