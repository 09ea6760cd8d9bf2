#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rlutrans_gpu.py -m gpu -q --no-header -rA 2>&1 | grep -E "max-abs|passed|failed|Error|error" | tee gpurun_out/rlutrans.log
timeout 300 python tools/bench_transblock.py 2>&1 | tail -2 | tee gpurun_out/rlutrans_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tb_ -c 12 python tools/bench_transblock.py 2>&1 | grep -E "tb_|gpu__time" | tail -12 | tee gpurun_out/rlutrans_ncu.log
