#!/bin/bash
# A/B of environment switches with the quick bench (sustained regime), interleaved twice: bash scripts/gpu_envab.sh "" "VAR=1" ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/envab.log
for rep in 1 2; do
  for v in "$@"; do
    env $v python bench.py --quick --no-side 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$v]', 'ms/pass %.4f' % d['ms_per_pass'], 'value %.0f' % d['value'], 'frac %.4f' % d['roofline']['frac'], d['clocks']['sm_mhz'], d['gpu_launches'])" | tee -a gpurun_out/envab.log
  done
done
