#!/bin/bash
# Runs on the GPU box: programmatic dependent launch - unit tests, parity tests, then A/B of the bench.
fail=0
for i in 0 1 2 3 4 5 6 7 8; do timeout 60 ./build/test_pair_tc_wd $i 3 > /tmp/p.log 2>&1 || { fail=1; tail -5 /tmp/p.log; }; done
for i in 0 3 5 6 11 12; do timeout 60 ./build/test_conv_tc_wd $i 3 > /tmp/c.log 2>&1 || { fail=1; tail -5 /tmp/c.log; }; done
echo "unit fail=$fail"; [ $fail = 0 ] || exit 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
run() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'ms/step %.3f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'])"; }
E2E_NO_PDL=1 run serial
run pdl
E2E_NO_PDL=1 run serial
run pdl
