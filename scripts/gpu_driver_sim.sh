#!/bin/bash
# what the driver does at round end, in one place: smoke(), the GPU suite, both bench arms with the driver's flags
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/sim_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/sim_smoke.log
( time python -m pytest tests/ -x -q -m gpu ) > gpurun_out/sim_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/sim_pytest.log
( time python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/sim_bench_ref.json 2> gpurun_out/sim_bench_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/sim_bench_ref.err
( time python3 bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/sim_bench.json 2> gpurun_out/sim_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/sim_bench.err
python - <<'PY'
import json
r=json.loads(open('gpurun_out/sim_bench_ref.json').read().strip().splitlines()[-1]); d=json.loads(open('gpurun_out/sim_bench.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['config']==({k:d['config'][k] for k in r['config']}), r['cpu_baseline']['kind'])
print('ours', d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['steps'], d['warmup'])
print('ratio e2e', d['e2e']['value']/r['e2e']['value'])
PY
