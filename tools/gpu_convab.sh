#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --no-header -rA -k "golden or oracle or ffconv or determin" 2>&1 | grep -E "passed|failed|Error|error" | tee gpurun_out/quick.log
echo "WGS=1"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python tools/show_bench.py
M2T_CONV_WGS=2 python -m m2trans_b200.build --force > /dev/null 2>&1
echo "WGS=2"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python tools/show_bench.py
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --no-header -rA -k "golden or oracle or ffconv or determin" 2>&1 | grep -E "passed|failed|Error|error"
