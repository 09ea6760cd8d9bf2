#!/bin/bash
mkdir -p gpurun_out
echo "== conv umma"; timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -rA --no-header -k "stage_ffconv" 2>&1 | tail -30 | tee gpurun_out/conv.log
echo "== engine"; timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --no-header -k "not stage_" 2>&1 | tail -30 | tee gpurun_out/engine.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench.log
