#!/bin/bash
# last GPU seconds of the round: LoRA vs full-training step time at the benchmark shape
mkdir -p gpurun_out
timeout 100 python tools/lora_step_time.py 5 > gpurun_out/s32_lora_time.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/s32_lora_time.log
