#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/pairwide.log; : > $L
for i in 5 6 8; do timeout 120 build/test_pair_tc_wd $i 1 >> $L 2>&1; echo "rc=$?" >> $L; done
for i in 14 15; do timeout 120 build/test_pair_tc $i 5 >> $L 2>&1; done
grep -E "PASS|FAIL|rc=|time|WATCHDOG" $L
timeout 600 python -m pytest tests/test_gpu_vocoder.py -q -x 2>&1 | tail -2
timeout 300 python bench.py --quick --no-side 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/pass %.4f' % d['ms_per_pass'], 'value %.0f' % d['value'], 'frac %.4f' % d['roofline']['frac'], d['clocks']['sm_mhz'])"
