#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -k "ffconv or golden or cfg2 or batch_consistency" > gpurun_out/r2i_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2i_tests.log; tail -4 gpurun_out/r2i_tests.log
timeout 300 python bench.py --no-cpu > gpurun_out/r2i_bench.log 2>&1; echo "cfg2: $(tail -1 gpurun_out/r2i_bench.log | python tools/show_bench.py)"
timeout 300 python tools/stage_profile.py cfg2 2>&1 | cut -c1-120 > gpurun_out/r2i_stage_cfg2.log; head -3 gpurun_out/r2i_stage_cfg2.log; tail -1 gpurun_out/r2i_stage_cfg2.log
timeout 300 python tools/stage_profile.py cfg4 2>&1 | cut -c1-120 > gpurun_out/r2i_stage_cfg4.log; head -3 gpurun_out/r2i_stage_cfg4.log; tail -1 gpurun_out/r2i_stage_cfg4.log
bash tools/gpu_conv_timing.sh
