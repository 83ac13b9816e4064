#!/bin/bash
# round-2 GPU session 1: parity groups, A/B tools, bench arms.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/s1_gpu.txt 2>&1
timeout 1500 python tools/run_gpu_checks.py > gpurun_out/s1_checks.log 2>&1; echo "checks rc=$?"
cp gpurun_out/checks.json gpurun_out/s1_checks.json
tail -5 gpurun_out/s1_checks.log
timeout 300 python tools/gemm_vs_cublas.py > gpurun_out/s1_gemm_vs_cublas.txt 2>&1; echo "gemm_vs_cublas rc=$?"
timeout 300 python tools/attn_vs_sdpa.py > gpurun_out/s1_attn_vs_sdpa.txt 2>&1; echo "attn_vs_sdpa rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference-gpu --steps 10 --warmup 3 > gpurun_out/s1_bench_refgpu.json 2> gpurun_out/s1_bench_refgpu.err; echo "refgpu rc=$?"
timeout 600 python bench.py --api dropin --steps 10 --warmup 3 --no-generate --no-cpu-baseline --no-hbm-kernels > gpurun_out/s1_bench_dropin.json 2> gpurun_out/s1_bench_dropin.err; echo "dropin rc=$?"
timeout 600 python bench.py --model tv2o-large --events 4096 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_large.json 2> gpurun_out/s1_bench_large.err; echo "large rc=$?"
timeout 600 python bench.py --impl reference-gpu --model tv2o-large --events 4096 --batch 8 --steps 3 --warmup 3 > gpurun_out/s1_bench_large_refgpu.json 2> gpurun_out/s1_bench_large_refgpu.err; echo "large refgpu rc=$?"
timeout 400 python tools/cpu_thread_sweep.py > gpurun_out/s1_cpu_sweep.log 2>&1; echo "sweep rc=$?"
for f in gpurun_out/s1_bench*.json; do echo "== $f"; head -c 600 $f; echo; done
