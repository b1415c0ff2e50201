#!/bin/bash
# Runs on the GPU box: the evidence committed under profiles/ for round 2 (scripts/collect_profiles_r2.py <tag> turns
# gpurun_out/ into profiles/r02_*):
#   1. pytest -m gpu (with the parity-margin log), the default bench line (all workloads + baselines), the reference arm
#   2. ncu launch list of one short bench run with per-launch DRAM bytes and tensor-pipe activity (time + traffic + the
#      BASELINE metric's "vocoder tensor-pipe util %" per launch)
#   3. ncu --set full of the dominant kernel (stage-1 k = 11 pair), the fused whole-resblock kernel, the stage-3 pair_tz
#      kernel (k = 11, d = 1) and the mel kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-a}
rm -f gpurun_out/parity_margins.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -n 2 gpurun_out/pytest_gpu_$TAG.log
cp gpurun_out/parity_margins.jsonl gpurun_out/parity_margins_$TAG.jsonl
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-300 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$TAG.json
E2E_OPERAND_DTYPE=fp16 timeout 600 python bench.py --quick --no-side > gpurun_out/bench_fp16_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_fp16_$TAG.json
CMD="python bench.py --steps 1 --warmup 3 --passes 1 --quick --no-side"
$CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second --clock-control none -c 520 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_tc -s 3 -c 1 -f -o gpurun_out/prof_pair_s1k11_$TAG $CMD > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rb_tc -s 1 -c 1 -f -o gpurun_out/prof_rb_s2k3_$TAG $CMD > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_tz -s 3 -c 1 -f -o gpurun_out/prof_tz_s3k11_$TAG $CMD > gpurun_out/ncu_f4.log 2>&1
python scripts/time_mel.py 1024 220500 3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mel_kernel -s 3 -c 1 -f -o gpurun_out/prof_mel_$TAG python scripts/time_mel.py 1024 220500 3 > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out/*_$TAG.* | head -20
