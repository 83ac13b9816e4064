#!/bin/bash
# session 14: (1) the concurrent generate_stream checks, (2) PDL on/off A/B on one box
mkdir -p gpurun_out
timeout 900 python tools/run_gpu_checks.py gemm gemm_exact model_train model_vs_hf model_peaked_greedy model_generate > gpurun_out/s14_checks.log 2>&1; echo "checks rc=$?"
grep -nE "ok in|FAIL|TOTAL|stream_" gpurun_out/s14_checks.log | head -40
for rep in 1 2; do
for pdl in 0 1; do
  B200_PDL=$pdl timeout 400 python bench.py --steps 12 --warmup 4 --no-hbm-kernels > gpurun_out/s14_bench_pdl${pdl}_$rep.json 2> gpurun_out/s14_bench_pdl${pdl}_$rep.err; echo "pdl=$pdl rep=$rep rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/s14_bench_pdl${pdl}_$rep.json'))
print('pdl=$pdl', d['ms_per_step'], 'gemm frac', d['roofline']['frac'], 'clk', d['clocks']['sm_mhz'], 'loss', d.get('loss'))
PY
done
done
