#!/bin/bash
mkdir -p gpurun_out
export B200_DECODE_PROFILE=
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_events --launch-skip 5 --launch-count 1 -o gpurun_out/s5_decode_b1 -f python tools/decode_profile.py 1 512 > gpurun_out/s5_ncu_b1.log 2>&1; tail -3 gpurun_out/s5_ncu_b1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_events --launch-skip 5 --launch-count 1 -o gpurun_out/s5_decode_b8 -f python tools/decode_profile.py 8 1024 > gpurun_out/s5_ncu_b8.log 2>&1; tail -3 gpurun_out/s5_ncu_b8.log
ls -la gpurun_out/s5_*.ncu-rep
timeout 300 python tools/run_gpu_checks.py loss_optim > gpurun_out/s5_checks.log 2>&1; grep -n "FAIL\|CRASH\|^\[\|TOTAL" gpurun_out/s5_checks.log
