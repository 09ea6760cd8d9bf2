#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2g_tests.log
tail -6 gpurun_out/r2g_tests.log
timeout 600 python bench.py > gpurun_out/r2g_bench.log 2>&1; tail -1 gpurun_out/r2g_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),d['e2e']['ms_per_step'],'wall',d['e2e']['wall_ms_per_step'],'e2e_fp32',round(d['e2e_fp32']['value'],1),'eval',d['e2e_eval'])
print('roofline',d['roofline']['kernel'][:40],d['roofline']['frac'],'second',d['roofline_second']['frac'],'fwd',d['roofline_forward']['frac'])
for k in d['kernels']: print(k)
print(d['largest_launch'], d.get('cpu_baseline'), d.get('parity'))
"
