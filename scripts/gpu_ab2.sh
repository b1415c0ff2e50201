#!/bin/bash
# A/B of in-tree library builds (E2E_TTS_B200_LIB) with the quick bench in the >= 2 s sustained regime, interleaved
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/ab2.log
for rep in 1 2; do
  for v in "$@"; do
    if [ "$v" = "base" ]; then L=$PWD/e2e_tts_b200/lib/libe2e_tts_b200.so; else L=$PWD/e2e_tts_b200/lib/ab_$v.so; fi
    E2E_TTS_B200_LIB=$L python bench.py --quick --no-side 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'ms/pass %.4f' % d['ms_per_pass'], 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'frac %.4f' % d['roofline']['frac'], d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab2.log
  done
done
