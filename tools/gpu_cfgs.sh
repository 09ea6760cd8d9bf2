#!/bin/bash
for w in cfg1 cfg3 cfg4 cfg2; do
  echo "== $w"; timeout 900 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'ffconv GB/s', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), 'fwd frac', round(d['roofline_forward']['frac'],4), 'cpu', d.get('cpu_baseline',{}).get('value'), 'parity', d.get('parity'))"
done
