#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/micro/mufu_bench 2>&1 | tee gpurun_out/s20_mufu_bench.txt
for v in 0 3; do
  echo "== VAR $v"; B200_ATTN_FWD_VAR=$v timeout 200 python tools/attn_fwd_profile.py 2>&1 | grep -vE "UserWarning" | head -10 | tee gpurun_out/s20_prof_var$v.txt
done
