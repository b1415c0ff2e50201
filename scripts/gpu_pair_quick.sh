#!/bin/bash
# Runs on the GPU box: pair-kernel unit tests (watchdog flavour, small shapes), then the perf shapes, then pytest + bench.
fail=0
for i in 0 1 2 3 4 5 6 7 8; do timeout 60 ./build/test_pair_tc_wd $i 3 > /tmp/p.log 2>&1 || { fail=1; grep -E "pair|mismatch|WATCHDOG|error|act:" /tmp/p.log | head -8; }; done
echo "unit fail=$fail"; [ $fail = 0 ] || exit 1
for i in 9 10 11 12 13 14 15; do timeout 120 ./build/test_pair_tc $i 20 | grep -E "perf|time|FAIL" | paste - - | sed 's/  */ /g' | cut -c1-140; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
for r in 1 2; do python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step %.3f value %.0f e2e %.0f frac %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac']))"; done
