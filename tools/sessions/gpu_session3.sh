#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/decode_profile.py 1 512 > gpurun_out/s3_decode_profile_b1.txt 2>&1; cat gpurun_out/s3_decode_profile_b1.txt | tail -20
timeout 300 python tools/decode_profile.py 8 2047 > gpurun_out/s3_decode_profile_b8.txt 2>&1; cat gpurun_out/s3_decode_profile_b8.txt | tail -20
timeout 900 python tools/run_gpu_checks.py loss_optim model_train model_medium_long > gpurun_out/s3_checks.log 2>&1
grep -n "FAIL\|CRASH\|^\[\|TOTAL\|medium peaked\|long:" gpurun_out/s3_checks.log | head -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rmsnorm|swiglu|rope_kernel|ce_fwd|ce_bwd|tiny_attn|colsum" --launch-skip 14 --launch-count 14 -o gpurun_out/s3_hbm_kernels -f python tools/hbm_once.py > gpurun_out/s3_ncu_hbm.log 2>&1; tail -3 gpurun_out/s3_ncu_hbm.log
ls -la gpurun_out/s3_hbm_kernels.ncu-rep
