#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --no-header -rA -k "fused_tail or golden or oracle" 2>&1 | grep -E "PSNR|max-abs|passed|failed|Error|error" | grep -v "variant=15" | tee gpurun_out/tail.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330 | tee gpurun_out/bench.log
