#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc05_v3 -s 1 -c 1 -f -o gpurun_out/s16_attn_v3 python tools/attn_fwd_once.py 3 > gpurun_out/s16_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/s16_ncu.log
