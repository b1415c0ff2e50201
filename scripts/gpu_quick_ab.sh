#!/bin/bash
# Runs on the GPU box: quick same-box A/B of environment switches.  usage: gpu_quick_ab.sh "VAR=val" "VAR2=val" ...
# ("-" = defaults).  Per setting: two bench runs (30 steps) and the ncu time of the kernels matching $KREGEX (optional).
mkdir -p gpurun_out
[ -n "$PYTEST" ] && { timeout 900 python -m pytest $PYTEST -m gpu -x -q 2>&1 | tail -n 2; }
[ -z "$NOBENCH" ] && for rep in 1 2; do
  for s in "$@"; do
    e=$s; [ "$s" = "-" ] && e="E2E_DUMMY=1"
    env $e timeout 180 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$s', 'ms/step %.3f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'])"
  done
done | tee gpurun_out/quick_ab.log
if [ -n "$KREGEX" ]; then
  for s in "$@"; do
    e=$s; [ "$s" = "-" ] && e="E2E_DUMMY=1"
    env $e ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$KREGEX -s ${KSKIP:-0} -c ${KCOUNT:-4} --csv --log-file gpurun_out/k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
    echo "$s: $(grep '^"' gpurun_out/k.csv | tail -n +2 | python -c "
import csv,sys
print(' '.join('%s=%.1fus' % (r[4].split('(')[0].split('<')[0][-24:], float(r[-1])/1e3) for r in csv.reader(sys.stdin)))")"
  done | tee gpurun_out/quick_ab_kernels.log
fi
