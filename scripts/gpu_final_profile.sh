#!/bin/bash
# Runs on the GPU box: the evidence committed under profiles/ for this round.
#   1. pytest -m gpu, default bench (with cpu_baseline), reference arm
#   2. ncu launch list of one bench run with per-launch DRAM bytes (time + traffic per launch)
#   3. ncu --set full of the dominant kernel (stage-1 k=11 pair) and of the mel kernel
#   4. config-3 (8 x 30 s), config-5 (mel, 1024 x 10 s) and iSTFTNet timings
mkdir -p gpurun_out
TAG=${1:-v8}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -n 2 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-300 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$TAG.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_tc -s 6 -c 1 -f -o gpurun_out/prof_pair_s1k11_$TAG $CMD > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 20 -c 1 -f -o gpurun_out/prof_conv_s0k11_$TAG $CMD > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mel_kernel -s 2 -c 1 -f -o gpurun_out/prof_mel_$TAG python scripts/time_mel.py 1024 220500 2 > gpurun_out/ncu_f3.log 2>&1
python scripts/time_voc.py 8 2584 10 2>&1 | tail -n 2 | tee gpurun_out/time_voc_cfg3_$TAG.log
python scripts/time_mel.py 1024 220500 20 | tee gpurun_out/time_mel_$TAG.log
python scripts/time_istft.py 16 431 20 | tee gpurun_out/time_istft_$TAG.log
