#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_.*_tc05 -c 2 -f -o gpurun_out/s30_attn_final python tools/attn_bwd_once.py 1 > gpurun_out/s30_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/s30_ncu.log
