#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -s -k "fused_tail or golden or oracle or cfg2" > gpurun_out/r2d_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2d_tests.log
grep -E "strip vs|passed|failed|rc=|Error|error" gpurun_out/r2d_tests.log | tail -16
timeout 300 python bench.py --no-cpu > gpurun_out/r2d_bench.log 2>&1; echo "cfg2: $(tail -1 gpurun_out/r2d_bench.log | python tools/show_bench.py)"
timeout 300 python tools/stage_profile.py cfg2 2>&1 | cut -c1-100 > gpurun_out/r2d_stage_cfg2.log; cat gpurun_out/r2d_stage_cfg2.log
timeout 300 python tools/stage_profile.py cfg4 2>&1 | cut -c1-100 > gpurun_out/r2d_stage_cfg4.log; cat gpurun_out/r2d_stage_cfg4.log
