#!/bin/bash
# session 26 (final commit): full GPU suite, bench, smoke, sanitizer over the attention kernels, launch list of one step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s26_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s26_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s26_bench.json 2> gpurun_out/s26_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s26_bench.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['clocks'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['roofline']['traffic'])
for k,v in d['hbm_kernels']['kernels'].items(): print('  ',k, v['us'], v['frac'])
g=d['generate']; print({k:(v['events_per_s'],v.get('graph_loop_events_per_s'),v['roofline']['frac']) for k,v in g.items() if k.startswith('batch')})
PY
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
# e2e vs device-resident A/B of the attention forward generations (the e2e leg syncs on the loss every step)
for gen in v3 v2 v3; do
  B200_ATTN_FWD_TC=$gen timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-generate --no-hbm-kernels > gpurun_out/s26_bench_$gen.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/s26_bench_$gen.json'));print('$gen', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
done
for t in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $t --error-exitcode 7 python tools/attn_small.py > gpurun_out/s26_${t}_attn.log 2>&1; echo "$t attn rc=$?"
  grep -E "SUMMARY|worst" gpurun_out/s26_${t}_attn.log | tail -3
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/s26_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-generate --no-hbm-kernels > gpurun_out/s26_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python tools/summarize_launches.py gpurun_out/s26_launches.csv gpurun_out/s26 | head -32
