#!/bin/bash
# session 15: third-generation attention forward (persistent, two query tiles per CTA): parity + timing vs v2 and SDPA
mkdir -p gpurun_out
timeout 300 python tools/run_gpu_checks.py attn_tc05 > gpurun_out/s15_attn_v3.log 2>&1; echo "v3 checks rc=$?"
grep -nE "ok in|FAIL|TOTAL|tc_fwd|tc_vs_mma_fwd|tc_vs_mma_lse|time_ms_tc|tflops_tc|timeout|error" gpurun_out/s15_attn_v3.log | head -40
B200_ATTN_FWD_TC=v2 timeout 300 python tools/run_gpu_checks.py attn_tc05 > gpurun_out/s15_attn_v2.log 2>&1; echo "v2 checks rc=$?"
grep -nE "time_ms_tc|tflops_tc" gpurun_out/s15_attn_v2.log
timeout 300 python tools/attn_vs_sdpa.py > gpurun_out/s15_attn_vs_sdpa.txt 2>&1; echo "sdpa rc=$?"; grep -E "event_fwd|event_bwd" gpurun_out/s15_attn_vs_sdpa.txt
