#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/decode_profile.py 1 512 > gpurun_out/s9_decode_profile_b1.txt 2>&1; tail -18 gpurun_out/s9_decode_profile_b1.txt
B200_DECODE_L2_PREFETCH=0 timeout 300 python tools/decode_profile.py 1 512 > gpurun_out/s9_decode_profile_b1_nol2.txt 2>&1; tail -18 gpurun_out/s9_decode_profile_b1_nol2.txt | head -3
timeout 300 python tools/decode_profile.py 8 2047 > gpurun_out/s9_decode_profile_b8.txt 2>&1; tail -18 gpurun_out/s9_decode_profile_b8.txt | head -3
B200_DECODE_L2_PREFETCH=0 timeout 300 python tools/decode_profile.py 8 2047 > gpurun_out/s9_decode_profile_b8_nol2.txt 2>&1; tail -18 gpurun_out/s9_decode_profile_b8_nol2.txt | head -3
timeout 900 python tools/run_gpu_checks.py decode model_generate model_peaked_greedy model_medium_long gemm_wgrad gemm_exact > gpurun_out/s9_checks.log 2>&1
grep -n "FAIL\|CRASH\|^\[\|TOTAL" gpurun_out/s9_checks.log | head -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s9_bench.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['clocks'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['whole_step_frac'])
for k,v in d['hbm_kernels']['kernels'].items(): print('  ',k, v['us'], v['frac'])
print(json.dumps(d['generate'])[:1200])
PY
