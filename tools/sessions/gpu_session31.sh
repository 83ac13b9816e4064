#!/bin/bash
# one short call (5.7 GPU-minutes were left): smoke, the LoRA group, the training parity group
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s31_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/s31_smoke.log
timeout 240 python tools/run_gpu_checks.py lora_train > gpurun_out/s31_lora.log 2>&1; echo "lora rc=$?"; cp gpurun_out/checks.json gpurun_out/s31_lora_checks.json 2>/dev/null; tail -60 gpurun_out/s31_lora.log
timeout 240 python tools/run_gpu_checks.py model_train > gpurun_out/s31_train.log 2>&1; echo "train rc=$?"; tail -40 gpurun_out/s31_train.log
