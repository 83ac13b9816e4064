#!/bin/bash
mkdir -p gpurun_out
# ---- CTA-pair GEMM: correctness, then A/B
timeout 900 python tools/run_gpu_checks.py gemm_fwd gemm_swiglu gemm_dgrad gemm_wgrad gemm_exact fused_rope model_train > gpurun_out/s8_checks_pair.log 2>&1
grep -n "FAIL\|CRASH\|^\[\|TOTAL" gpurun_out/s8_checks_pair.log | head -20
grep -n "mbarrier wait timeout\|Error\|error" gpurun_out/s8_checks_pair.log | head -5
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-generate --no-cpu-baseline --no-hbm-kernels > gpurun_out/s8_bench_$name.json 2> gpurun_out/s8_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8_bench_$name.json"))
    print("$name", round(d["ms_per_step"],3), "ms/step  gemm", round(d["roofline"]["gemm_ms_per_step"],2), "ms  frac", round(d["roofline"]["frac"],3), "clocks", d["clocks"]["sm_mhz"], "loss", d["e2e"]["last_loss"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/s8_bench_$name.err").read()[-600:])
PY
}
run pair X=1
run nopair B200_GEMM_CTA_PAIR=0
run pair2 X=1
timeout 300 python tools/gemm_vs_cublas.py > gpurun_out/s8_gemm_vs_cublas_pair.txt 2>&1; tail -3 gpurun_out/s8_gemm_vs_cublas_pair.txt
# ---- persistent decode after the code-size restructure + L2 bulk prefetch
timeout 300 python tools/decode_profile.py 1 512 > gpurun_out/s8_decode_profile_b1.txt 2>&1; tail -18 gpurun_out/s8_decode_profile_b1.txt
timeout 300 python tools/decode_profile.py 8 2047 > gpurun_out/s8_decode_profile_b8.txt 2>&1; tail -18 gpurun_out/s8_decode_profile_b8.txt
timeout 900 python tools/run_gpu_checks.py decode model_generate model_peaked_greedy model_medium_long > gpurun_out/s8_checks_decode.log 2>&1
grep -n "FAIL\|CRASH\|^\[\|TOTAL" gpurun_out/s8_checks_decode.log | head -20
