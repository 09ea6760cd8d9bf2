#!/bin/bash
# multi-GPU lines: usage gpu_r2_multi.sh N  (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
run() { # workload impl
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $1 --no-cpu > gpurun_out/r2m_n${N}_$1.log 2>&1
  tail -1 gpurun_out/r2m_n${N}_$1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1 N=$N', d['scaling'], 'value', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],3), 'e2e', round(d.get('e2e',{}).get('value',0),1), 'e2e_fp32', round(d.get('e2e_fp32',{}).get('value',0),1), 'eval', round(d.get('e2e_eval',{}).get('value',0),1), d.get('clocks',{}).get('reasons'))
" || tail -5 gpurun_out/r2m_n${N}_$1.log
}
run cfg2; run cfg3; run cfg4; run cfg5
