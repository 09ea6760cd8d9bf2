#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_clip_gpu.py tests/test_engine_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2w_tests.log
timeout 300 python tools/bench_clip.py --batch 256 2>&1 | tail -2 | cut -c1-300 | tee gpurun_out/r2w_clip256.log
timeout 300 python tools/bench_clip.py --batch 32 2>&1 | tail -2 | cut -c1-300 | tee gpurun_out/r2w_clip32.log
timeout 300 python bench.py --workload cfg5 --no-cpu 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r2w_cfg5.log
