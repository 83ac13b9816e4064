#!/bin/bash
mkdir -p gpurun_out
echo "== pingpong on"; timeout 200 python tools/attn_fwd_profile.py 2>&1 | tee gpurun_out/s18_prof_pp1.txt
echo "== pingpong off"; B200_ATTN_FWD_PINGPONG=0 timeout 200 python tools/attn_fwd_profile.py 2>&1 | tee gpurun_out/s18_prof_pp0.txt
