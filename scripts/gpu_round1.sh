#!/bin/bash
# Runs on the GPU box: pytest -m gpu, bench, traced pair-kernel perf shapes, ncu launch list + one --set full capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -3 gpurun_out/pytest_gpu_all.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
for i in 9 11 12 13 14 15; do timeout 120 ./build/test_pair_tc_trace $i 20; done > gpurun_out/pair_trace.log 2>&1
E2E_PAIR_CG=1 timeout 120 ./build/test_pair_tc_trace 9 20 >> gpurun_out/pair_trace.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_v7.csv $CMD > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_tc -s 6 -c 1 -f -o gpurun_out/prof_pair_s1k11 $CMD > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_tc -s 18 -c 1 -f -o gpurun_out/prof_pair_s3k3 $CMD > gpurun_out/ncu_f2.log 2>&1
tail -2 gpurun_out/ncu_l.log gpurun_out/ncu_f1.log gpurun_out/ncu_f2.log
