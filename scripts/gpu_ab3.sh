#!/bin/bash
# packed-fp32 epilogue math: device units of all three fused kernels, vocoder parity, same-box A/B of three library builds
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/ab3_units.log
: > $L
for i in 0 1 2 3 4 5 6 7 8; do timeout 120 build/test_pair_tc_wd $i 1 >> $L 2>&1; timeout 120 build/test_rb_tc_wd $i 1 >> $L 2>&1; done
for i in 0 3 6 9 11; do timeout 120 build/test_pair_tz_wd $i 1 >> $L 2>&1; done
for i in 0 1 2 3 4 5 6 10 11 12 13 14; do timeout 120 build/test_conv_tc_wd $i 1 >> $L 2>&1; done
echo "units: $(grep -c PASS $L) pass, $(grep -c FAIL $L) fail"; grep -E "FAIL|WATCHDOG|mismatch" $L | head
timeout 1200 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_full_size.py tests/test_postnet.py tests/test_gpu_istft.py -x -q -m gpu > gpurun_out/ab3_pytest.log 2>&1; tail -3 gpurun_out/ab3_pytest.log
bash scripts/gpu_ab.sh old scalar new
for rep in 1 2; do E2E_TZ_K3=1 E2E_TTS_B200_LIB=$PWD/e2e_tts_b200/lib/ab_new.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new+TZ_K3', 'ms/step %.3f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'])"; done | tee -a gpurun_out/ab.log
