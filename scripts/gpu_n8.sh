#!/bin/bash
# N GPUs (default 8): the cfg4 bench line (strong scaling, timed + verified gather)
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/n${N}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/n${N}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_pass','scaling','n_gpus')}); print(d['e2e']); print(d.get('gather')); print(d.get('weak16')); print(d['clocks'])
PY
