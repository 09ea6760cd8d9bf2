#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -k "fused_tail or golden" > gpurun_out/r2f_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2f_tests.log; tail -3 gpurun_out/r2f_tests.log
timeout 300 python bench.py --no-cpu > gpurun_out/r2f_bench.log 2>&1; echo "cfg2: $(tail -1 gpurun_out/r2f_bench.log | python tools/show_bench.py)"
timeout 300 python tools/stage_profile.py cfg2 2>&1 | cut -c1-100 > gpurun_out/r2f_stage_cfg2.log; grep -E "tail|total" gpurun_out/r2f_stage_cfg2.log
timeout 300 python tools/stage_profile.py cfg1 2>&1 | cut -c1-100 > gpurun_out/r2f_stage_cfg1.log; cat gpurun_out/r2f_stage_cfg1.log
