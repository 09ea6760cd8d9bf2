#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q --no-header -rA 2>&1 | grep -E "PSNR|passed|failed|Error|error" | grep -v "variant=15" | tee gpurun_out/all.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330 | tee gpurun_out/bench.log
