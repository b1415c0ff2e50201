#!/bin/bash
# session-3 call A: mel A/B + plan dump + per-launch conv_tc times with the staged-store switch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_mel_ab.sh
E2E_DUMP_PLAN=1 timeout 300 python scripts/time_voc.py 16 431 3 2> gpurun_out/plan_dump.log | tail -2
grep -c plan gpurun_out/plan_dump.log
NOBENCH=1 KREGEX=conv_tc KCOUNT=23 bash scripts/gpu_quick_ab.sh - E2E_CONV_STAGED=1 E2E_CONV_STAGED=0
