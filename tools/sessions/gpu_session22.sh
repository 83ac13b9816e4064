#!/bin/bash
# session 22: attention forward v3 as default + backward without the staging barrier: parity, timing, full GPU suite, bench
mkdir -p gpurun_out
timeout 300 python tools/run_gpu_checks.py attn_tc05 > gpurun_out/s22_attn.log 2>&1; echo "attn checks rc=$?"
grep -nE "ok in|FAIL|TOTAL|time_ms|tflops_tc|timeout|rror" gpurun_out/s22_attn.log | head -20
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s22_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s22_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s22_bench.json 2> gpurun_out/s22_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s22_bench.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['clocks'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['whole_step_frac'])
g=d['generate']; print({k:(v['events_per_s'],v.get('graph_loop_events_per_s'),v['roofline']['frac']) for k,v in g.items() if k.startswith('batch')})
PY
