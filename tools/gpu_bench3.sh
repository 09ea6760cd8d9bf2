#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv | tee gpurun_out/smi.txt
for i in 1 2 3; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'], d['e2e']['value'])"; done | tee gpurun_out/bench3.log
