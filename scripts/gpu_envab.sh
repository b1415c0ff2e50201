#!/bin/bash
# Runs on the GPU box: pytest -m gpu, then A/B of the bench with / without an environment switch ($1, e.g. E2E_NO_TILED_SUMS).
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
run() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'ms/step %.3f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'])"; }
for r in 1 2; do env $1=1 bash -c "$(declare -f run); run with_$1"; run default; done
