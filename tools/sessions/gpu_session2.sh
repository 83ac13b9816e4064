#!/bin/bash
# round-2 GPU session 2: fixed parity groups + persistent decode kernel + compute-sanitizer
mkdir -p gpurun_out
timeout 1500 python tools/run_gpu_checks.py decode_paged model_generate model_peaked_greedy model_medium_long model_vs_hf gemm_exact decode > gpurun_out/s2_checks.log 2>&1; echo "checks rc=$?"
cp gpurun_out/checks.json gpurun_out/s2_checks.json
grep -n "FAIL\|CRASH\|^\[\|TOTAL" gpurun_out/s2_checks.log | head -40
timeout 300 python tools/attn_vs_sdpa.py > gpurun_out/s2_attn_vs_sdpa.txt 2>&1; echo "attn_vs_sdpa rc=$?"; tail -6 gpurun_out/s2_attn_vs_sdpa.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hbm-kernels > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/s2_bench.json'))
    print(d['ms_per_step'], json.dumps(d.get('generate'))[:1500])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/s2_bench.err').read()[-2000:])
PY
# compute-sanitizer: memcheck over the kernel-level groups and the generate loops (tiny shapes), racecheck over the shared-memory kernels
for g in gemm_fwd gemm_dgrad gemm_wgrad elementwise attn_tiny loss_optim decode decode_paged model_generate; do
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/run_gpu_checks.py --child $g > gpurun_out/s2_memcheck_$g.log 2>&1; echo "memcheck $g rc=$?"
  grep -c "ERROR SUMMARY" gpurun_out/s2_memcheck_$g.log; grep "ERROR SUMMARY" gpurun_out/s2_memcheck_$g.log | tail -1
done
for g in elementwise attn_tiny loss_optim decode; do
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/run_gpu_checks.py --child $g > gpurun_out/s2_racecheck_$g.log 2>&1; echo "racecheck $g rc=$?"
  grep "RACECHECK SUMMARY\|ERROR SUMMARY" gpurun_out/s2_racecheck_$g.log | tail -1
done
