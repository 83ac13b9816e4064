#!/bin/bash
mkdir -p gpurun_out
for u in 4 2; do
B200_ATTN_BWD_UNITS=$u timeout 300 python tools/run_gpu_checks.py attn_tc05 > gpurun_out/s24_attn_u$u.log 2>&1; echo "units=$u checks rc=$?"
grep -nE "ok in|FAIL|TOTAL|time_ms_bwd_tc|time_ms_tc |timeout|rror" gpurun_out/s24_attn_u$u.log | head -12
done
