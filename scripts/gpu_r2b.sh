#!/bin/bash
# round 2: fp16 operand mode + rb default rule: full GPU suite, then bench (quick) in both modes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -8 gpurun_out/r2b_pytest.log
timeout 300 python bench.py --quick --no-side > gpurun_out/r2b_bench_bf16.json 2> gpurun_out/r2b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_bf16.json')); print('bf16', d['ms_per_pass'], d['value'], d['roofline']['frac'])"
E2E_OPERAND_DTYPE=fp16 timeout 300 python bench.py --quick --no-side > gpurun_out/r2b_bench_fp16.json 2>> gpurun_out/r2b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_fp16.json')); print('fp16', d['ms_per_pass'], d['value'], d['roofline']['frac'])"
E2E_RB_FUSION=0 timeout 300 python bench.py --quick --no-side > gpurun_out/r2b_bench_norb.json 2>> gpurun_out/r2b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_norb.json')); print('bf16 no rb fusion', d['ms_per_pass'], d['value'], d['roofline']['frac'])"
