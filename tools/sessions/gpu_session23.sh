#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gemm_plan_sweep.py 2>&1 | grep -v Warning | tee gpurun_out/s23_gemm_plan_sweep.txt
