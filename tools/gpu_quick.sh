#!/bin/bash
# quick check: golden/oracle parity of the default path + 3 bench repeats
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --no-header -rA -k "golden or oracle or determin or cfg2" 2>&1 | grep -E "PSNR|passed|failed|Error|error" | grep -v "variant=15" | tee gpurun_out/quick.log
for i in 1 2 3; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"; done | tee gpurun_out/bench3.log
