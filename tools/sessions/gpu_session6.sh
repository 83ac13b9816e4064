#!/bin/bash
# A/B of the fused GEMM epilogues (step time, ms) + correctness of the 8-warp epilogue
mkdir -p gpurun_out
timeout 600 python tools/run_gpu_checks.py gemm_fwd gemm_swiglu gemm_dgrad gemm_wgrad fused_rope elementwise gemm_exact > gpurun_out/s6_checks.log 2>&1
grep -n "FAIL\|CRASH\|^\[\|TOTAL" gpurun_out/s6_checks.log
B200_GEMM_EG_PLAIN=2 timeout 300 python tools/run_gpu_checks.py gemm_fwd gemm_dgrad gemm_wgrad > gpurun_out/s6_checks_eg2.log 2>&1
grep -n "FAIL\|CRASH\|^\[\|TOTAL" gpurun_out/s6_checks_eg2.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-generate --no-cpu-baseline --no-hbm-kernels > gpurun_out/s6_bench_$name.json 2> gpurun_out/s6_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s6_bench_$name.json"))
    print("$name", round(d["ms_per_step"],3), "ms/step  gemm", round(d["roofline"]["gemm_ms_per_step"],2), "ms  clocks", d["clocks"]["sm_mhz"], "loss", d["e2e"]["last_loss"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/s6_bench_$name.err").read()[-800:])
PY
}
run base X=1
run swiglu8 B200_FUSE_SWIGLU=1
run swiglu4 B200_FUSE_SWIGLU=1 B200_GEMM_EG_FUSED=1
run rope8 B200_FUSE_ROPE_FWD=1
run both8 B200_FUSE_SWIGLU=1 B200_FUSE_ROPE_FWD=1
run plain8 B200_GEMM_EG_PLAIN=2
run base2 X=1
