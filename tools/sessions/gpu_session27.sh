#!/bin/bash
mkdir -p gpurun_out
for t in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $t --error-exitcode 7 python tools/attn_small.py > gpurun_out/s27_${t}_attn.log 2>&1; echo "$t attn rc=$?"
  grep -E "SUMMARY|worst|rel " gpurun_out/s27_${t}_attn.log | tail -5
done
