#!/bin/bash
mkdir -p gpurun_out
for m in tma split red; do
B200_ATTN_BWD_DQ=$m timeout 120 python tools/attn_bwd_once.py 3 2>&1 | grep "^mode"
done
for m in tma split; do
B200_ATTN_BWD_DQ=$m timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_sectors.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 6 -c 6 --csv --log-file gpurun_out/s25_bwd_$m.csv python tools/attn_bwd_once.py 1 > gpurun_out/s25_ncu_$m.log 2>&1; echo "ncu $m rc=$?"
done
python - <<'PY'
import csv
for m in ("tma","split"):
    rows=[r for r in csv.reader(open(f"gpurun_out/s25_bwd_{m}.csv")) if len(r)>10]
    hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
    agg={}
    for r in rows[1:]:
        k=(r[ix['ID']], r[ix['Kernel Name']][:50])
        agg.setdefault(k,{})[r[ix['Metric Name']]]=r[ix['Metric Value']]
    print("==",m)
    for k,v in agg.items(): print(k, v)
PY
