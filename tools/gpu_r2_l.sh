#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_metrics_gpu.py -x -q -s -k gmsd 2>&1 | tail -8
timeout 300 python tools/stage_profile.py cfg3 2>&1 | cut -c1-130 > gpurun_out/r2l_stage_cfg3.log; cat gpurun_out/r2l_stage_cfg3.log
timeout 300 python tools/stage_profile.py cfg1 2>&1 | cut -c1-130 > gpurun_out/r2l_stage_cfg1.log; cat gpurun_out/r2l_stage_cfg1.log
