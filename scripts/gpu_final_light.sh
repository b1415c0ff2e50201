#!/bin/bash
# Runs on the GPU box: scripts/gpu_final_profile.sh without the `ncu --set full` captures (kernels unchanged since the
# last ones): pytest -m gpu, smoke, default bench, reference arm, ncu launch list with DRAM bytes, other-config timings.
mkdir -p gpurun_out
TAG=${1:-v13}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -n 2 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 1
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-300 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$TAG.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l.log 2>&1
python scripts/time_voc.py 8 2584 10 2>&1 | tail -n 2 | tee gpurun_out/time_voc_cfg3_$TAG.log
python scripts/time_mel.py 1024 220500 20 | tee gpurun_out/time_mel_$TAG.log
python scripts/time_istft.py 16 431 20 | tee gpurun_out/time_istft_$TAG.log
python scripts/time_postnet.py 2>/dev/null | tee gpurun_out/time_postnet_$TAG.log
