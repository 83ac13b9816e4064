#!/bin/bash
# final commit: full GPU suite, default bench, smoke
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s29_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s29_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s29_bench.json 2> gpurun_out/s29_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s29_bench.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['roofline']['traffic'])
g=d['generate']; print({k:(v['events_per_s'],v.get('graph_loop_events_per_s'),v['roofline']['frac']) for k,v in g.items() if k.startswith('batch')})
print('cpu_baseline', d['cpu_baseline'])
PY
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
