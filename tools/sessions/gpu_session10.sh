#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s10_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s10_pytest_gpu.log
timeout 300 python tools/decode_profile.py 1 512 > gpurun_out/s10_decode_profile_b1.txt 2>&1; tail -18 gpurun_out/s10_decode_profile_b1.txt | head -3
timeout 300 python tools/decode_profile.py 8 2047 > gpurun_out/s10_decode_profile_b8.txt 2>&1; tail -18 gpurun_out/s10_decode_profile_b8.txt | head -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s10_bench.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['clocks'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['whole_step_frac'])
for k,v in d['hbm_kernels']['kernels'].items(): print('  ',k, v['us'], v['frac'])
g=d['generate']; print({k:(v['events_per_s'],v.get('graph_loop_events_per_s'),v['roofline']['frac']) for k,v in g.items() if k.startswith('batch')}, g.get('cpu_baseline'))
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/s10_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-generate --no-hbm-kernels > gpurun_out/s10_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python tools/summarize_launches.py gpurun_out/s10_launches.csv gpurun_out/s10 | head -40
timeout 600 python bench.py --model tv2o-large --events 4096 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s10_bench_large.json 2> gpurun_out/s10_bench_large.err; echo "large rc=$?"; head -c 400 gpurun_out/s10_bench_large.json; echo
timeout 600 python bench.py --api dropin --steps 10 --warmup 3 --no-generate --no-cpu-baseline --no-hbm-kernels > gpurun_out/s10_bench_dropin.json 2> gpurun_out/s10_bench_dropin.err; echo "dropin rc=$?"; head -c 300 gpurun_out/s10_bench_dropin.json; echo
