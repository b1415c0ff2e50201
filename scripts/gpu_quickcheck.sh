#!/bin/bash
# quick regression after a kernel change: device unit tests of one family + vocoder parity + quick bench (bf16 and fp16)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_zz_device_units.py tests/test_gpu_istft.py tests/test_postnet.py -q -x 2>&1 | tail -2
for rep in 1 2; do
timeout 300 python bench.py --quick --no-side 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bf16 ms/pass %.4f' % d['ms_per_pass'], 'value %.0f' % d['value'], 'frac %.4f' % d['roofline']['frac'], d['clocks']['sm_mhz'])"
E2E_OPERAND_DTYPE=fp16 timeout 300 python bench.py --quick --no-side 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fp16 ms/pass %.4f' % d['ms_per_pass'], 'value %.0f' % d['value'], 'frac %.4f' % d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
