#!/bin/bash
# Runs on the GPU box: A/B of in-tree library builds (E2E_TTS_B200_LIB) with the bench, then launch lists.
mkdir -p gpurun_out
for v in "$@"; do
  for rep in 1 2; do
    E2E_TTS_B200_LIB=$PWD/e2e_tts_b200/lib/ab_$v.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'ms/step %.3f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'])"
  done
done | tee gpurun_out/ab.log
[ -n "$NCU" ] && for v in "$@"; do
  E2E_TTS_B200_LIB=$PWD/e2e_tts_b200/lib/ab_$v.so ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_ab_$v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
done
