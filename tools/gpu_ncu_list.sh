#!/bin/bash
# launch list of one bench step (cold-cache, serialised: compare shares)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 160 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5]
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    name=d['Kernel Name'][:60]; v=float(d['Metric Value'].replace(',',''))
    unit=d['Metric Unit']
    if unit=='ns': v/=1e3
    elif unit=='ms': v*=1e3
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
with open('gpurun_out/launch_summary.txt','w') as f:
    for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        line=f"{t:10.1f} us {100*t/tot:5.1f}%  x{n:<4d} {k}"
        print(line); f.write(line+"\n")
    f.write(f"total {tot:.1f} us\n"); print("total",tot)
PY
