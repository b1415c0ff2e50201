#!/bin/bash
# round 2, first GPU pass: the changed GPU tests, the new bench line (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "not device_units" > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"
tail -3 gpurun_out/r2a_pytest.log
