#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python tools/show_bench.py; done | tee gpurun_out/bench3.log
bash tools/gpu_ncu_list.sh
