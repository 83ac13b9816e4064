#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/run_gpu_checks.py attn_tc05 > gpurun_out/s21_attn_v4.log 2>&1; echo "v4 checks rc=$?"
grep -nE "ok in|FAIL|TOTAL|time_ms_tc |tflops_tc|timeout|rror" gpurun_out/s21_attn_v4.log | head
echo "== v4 profile"; timeout 200 python tools/attn_fwd_profile.py 2>&1 | grep -vE "UserWarning" | head -20 | tee gpurun_out/s21_prof_v4.txt
echo "== v3 time"; B200_ATTN_FWD_TC=v3 timeout 200 python tools/attn_fwd_profile.py 2>&1 | grep "plain kernel"
