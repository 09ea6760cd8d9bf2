#!/bin/bash
mkdir -p gpurun_out
echo "== attn umma"; timeout 600 python -m pytest tests/test_stage_attn_gpu.py -m gpu -q -rA --no-header 2>&1 | grep -E "^C=|passed|failed|Error|error" | tee gpurun_out/attn.log
echo "== engine"; timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --no-header 2>&1 | tail -30 | tee gpurun_out/engine.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench.log
