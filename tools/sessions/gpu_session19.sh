#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 3 5 7; do
  echo "== VAR $v"; B200_ATTN_FWD_VAR=$v timeout 200 python tools/attn_fwd_profile.py 2>&1 | grep -vE "warp-tiles|^group 1|UserWarning" | head -9 | tee gpurun_out/s19_prof_var$v.txt
done
for v in 7 3; do
B200_ATTN_FWD_VAR=$v timeout 300 python tools/run_gpu_checks.py attn_tc05 > gpurun_out/s19_attn_var$v.log 2>&1; echo "VAR $v checks rc=$?"
grep -nE "ok in|FAIL|TOTAL|time_ms_tc |tflops_tc" gpurun_out/s19_attn_var$v.log
done
