#!/bin/bash
# round 2: graphs + N split + crop/mel: full GPU suite, then the full bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -8 gpurun_out/r2c_pytest.log
timeout 600 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print(d['ms_per_pass'], d['value'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['workloads'].items(): print(k, {a:b for a,b in v.items() if isinstance(b,(int,float))})
PY
E2E_NO_GRAPH=1 E2E_NO_NSPLIT=1 timeout 300 python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import e2e_tts_b200 as pkg
from e2e_tts_b200 import synthetic as sy
import bench
voc = pkg.HifiGan(sy.DEFAULT_CONFIG); voc.load_state_dict(sy.make_state_dict(sy.DEFAULT_CONFIG, 1, "strong")); voc = voc.eval().cuda()
for b in (1, 2, 4):
    mels = [sy.mel_like(b, 431, 500 + i).cuda() for i in range(4)]
    print("no graph, no N split: B=%d device ms %.4f" % (b, bench.time_forward(voc, mels, 100)))
PY
