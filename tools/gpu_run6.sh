#!/bin/bash
mkdir -p gpurun_out
echo "== stage tests"; timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_stage_attn_gpu.py -m gpu -q --no-header -k "stage_" 2>&1 | tail -5 | tee gpurun_out/stage.log
echo "== engine"; timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --no-header -rA -k "not stage_" 2>&1 | grep -E "PSNR|passed|failed|Error|error|rel err" | tee gpurun_out/engine.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -3 | cut -c1-400 | tee gpurun_out/bench.log
