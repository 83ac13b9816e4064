#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s13_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s13_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s13_bench.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['clocks'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['roofline']['traffic'])
for k,v in d['hbm_kernels']['kernels'].items(): print('  ',k, v['us'], v['frac'])
g=d['generate']; print({k:(v['events_per_s'],v.get('graph_loop_events_per_s'),v['roofline']['frac']) for k,v in g.items() if k.startswith('batch')})
PY
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
