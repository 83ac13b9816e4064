#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/run_gpu_checks.py attn_tc05 > gpurun_out/s17_attn_pp1.log 2>&1; echo "pingpong=1 checks rc=$?"
grep -nE "ok in|FAIL|TOTAL|time_ms_tc|tflops_tc|timeout|error" gpurun_out/s17_attn_pp1.log | head -20
B200_ATTN_FWD_PINGPONG=0 timeout 300 python tools/run_gpu_checks.py attn_tc05 > gpurun_out/s17_attn_pp0.log 2>&1; echo "pingpong=0 checks rc=$?"
grep -nE "ok in|FAIL|TOTAL|time_ms_tc |tflops_tc" gpurun_out/s17_attn_pp0.log
