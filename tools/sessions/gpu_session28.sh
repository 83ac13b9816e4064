#!/bin/bash
mkdir -p gpurun_out
for m in warp writer direct writer warp; do
  B200_ATTN_BWD_LSE=$m timeout 120 python tools/attn_bwd_once.py 3 2>&1 | grep "^mode" | sed "s/^mode/lse=$m mode/"
done
for m in writer direct; do
  B200_ATTN_BWD_LSE=$m timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/attn_small.py > gpurun_out/s28_racecheck_$m.log 2>&1; echo "racecheck $m rc=$?"
  grep -E "SUMMARY|worst" gpurun_out/s28_racecheck_$m.log | tail -2
done
B200_ATTN_BWD_LSE=writer timeout 300 python tools/run_gpu_checks.py attn_tc05 > gpurun_out/s28_attn_writer.log 2>&1; echo "attn checks (writer) rc=$?"
grep -nE "ok in|FAIL|TOTAL|time_ms_bwd_tc" gpurun_out/s28_attn_writer.log | head
