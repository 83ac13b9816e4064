#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/decode_profile.py 1 512 > gpurun_out/s4_decode_profile_b1.txt 2>&1; tail -22 gpurun_out/s4_decode_profile_b1.txt
timeout 300 python tools/decode_profile.py 8 2047 > gpurun_out/s4_decode_profile_b8.txt 2>&1; tail -22 gpurun_out/s4_decode_profile_b8.txt
timeout 900 python tools/run_gpu_checks.py decode model_generate model_peaked_greedy model_medium_long loss_optim > gpurun_out/s4_checks.log 2>&1
grep -n "FAIL\|CRASH\|^\[\|TOTAL" gpurun_out/s4_checks.log | head -30
