#!/bin/bash
# Runs on the GPU box: conv_tc standalone tests (watchdog flavour on the small shapes, then every configuration, direct
# and staged epilogue), pytest -m gpu, two bench runs.
mkdir -p gpurun_out
fail=0
: > gpurun_out/conv_all.log
n=$(./build/test_conv_tc list)
for st in 0 1; do for i in 0 1 2 3 4 5 6 10 11 12 13 14; do E2E_CONV_STAGED=$st timeout 60 ./build/test_conv_tc_wd $i 3 >> gpurun_out/conv_all.log 2>&1 || { fail=1; echo "wd cfg $i staged=$st FAILED"; }; done; done
[ $fail = 0 ] && for i in $(seq 0 $((n-1))); do timeout 120 ./build/test_conv_tc $i 20 >> gpurun_out/conv_all.log 2>&1 || { fail=1; echo "cfg $i FAILED"; }; done
grep -E "perf" -A6 gpurun_out/conv_all.log | grep -E "perf|time" | paste - - | sed 's/  */ /g' | cut -c1-140
echo "unit fail=$fail"
[ $fail = 0 ] || exit 1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -n 2 gpurun_out/pytest_gpu_all.log
for r in 1 2; do python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step %.3f' % d['ms_per_step'], 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'])"; done
