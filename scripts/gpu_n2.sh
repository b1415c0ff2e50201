#!/bin/bash
# two GPUs: the NCCL sharded-synthesis test, then the cfg4 bench line (strong scaling, timed + verified gather)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parallel_nccl.py -q -x > gpurun_out/n2_pytest.log 2>&1; echo "nccl pytest rc=$?"; tail -4 gpurun_out/n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/n2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n2_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_pass','scaling','n_gpus')}); print(d['e2e']); print(d.get('gather')); print(d.get('weak16')); print(d['config']['workload'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/n2_bench_ref.json 2>> gpurun_out/n2_bench.err; echo "ref rc=$?"; head -c 400 gpurun_out/n2_bench_ref.json
