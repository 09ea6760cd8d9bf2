#!/bin/bash
# round 2, first look at attn_z: stage tests, parity tests, per-kernel profile, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_stage_attnz_gpu.py -x -q -s > gpurun_out/r2a_attnz.log 2>&1
echo "attnz rc=$?" >> gpurun_out/r2a_attnz.log
timeout 900 python -m pytest tests/test_engine_gpu.py -q -s -k "attn_z or golden or oracle or speckle or precise" > gpurun_out/r2a_engine.log 2>&1
echo "engine rc=$?" >> gpurun_out/r2a_engine.log
timeout 300 python tools/stage_profile.py cfg2 > gpurun_out/r2a_stage_cfg2.log 2>&1
timeout 300 python tools/stage_profile.py cfg4 > gpurun_out/r2a_stage_cfg4.log 2>&1
timeout 300 python tools/stage_profile.py cfg1 > gpurun_out/r2a_stage_cfg1.log 2>&1
timeout 600 python bench.py --no-cpu > gpurun_out/r2a_bench.log 2>&1
tail -3 gpurun_out/r2a_attnz.log gpurun_out/r2a_engine.log
tail -12 gpurun_out/r2a_stage_cfg2.log
tail -1 gpurun_out/r2a_bench.log | python tools/show_bench.py || tail -c 600 gpurun_out/r2a_bench.log
