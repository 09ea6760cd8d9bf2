#!/bin/bash
mkdir -p gpurun_out
python tools/attn_timing.py 2>&1 | tail -24 | cut -c1-190
echo "== engine"; timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --no-header -rA -k "not stage_" 2>&1 | grep -E "PSNR|passed|failed|Error|error|rel err" | grep -v "variant=15" | tee gpurun_out/engine.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330 | tee gpurun_out/bench.log
