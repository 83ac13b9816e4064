#!/bin/bash
mkdir -p gpurun_out
./tools/micro/gridbar_bench > gpurun_out/s7_gridbar.txt 2>&1; cat gpurun_out/s7_gridbar.txt
timeout 300 python tools/decode_profile.py 1 512 > gpurun_out/s7_decode_profile_b1.txt 2>&1; tail -18 gpurun_out/s7_decode_profile_b1.txt
timeout 300 python tools/decode_profile.py 8 2047 > gpurun_out/s7_decode_profile_b8.txt 2>&1; tail -18 gpurun_out/s7_decode_profile_b8.txt
timeout 900 python tools/run_gpu_checks.py decode model_generate model_peaked_greedy model_medium_long > gpurun_out/s7_checks.log 2>&1
grep -n "FAIL\|CRASH\|^\[\|TOTAL" gpurun_out/s7_checks.log | head -30
