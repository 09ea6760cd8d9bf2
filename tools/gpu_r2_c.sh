#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stage_attnz_gpu.py tests/test_engine_gpu.py -x -q -k "attn_z or attnz or oracle or golden" > gpurun_out/r2c_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2c_tests.log
tail -3 gpurun_out/r2c_tests.log
timeout 300 python bench.py --no-cpu > gpurun_out/r2c_bench.log 2>&1; echo "cfg2: $(tail -1 gpurun_out/r2c_bench.log | python tools/show_bench.py)"
timeout 300 python bench.py --no-cpu --workload cfg1 > gpurun_out/r2c_cfg1.log 2>&1; echo "cfg1: $(tail -1 gpurun_out/r2c_cfg1.log | python tools/show_bench.py)"
timeout 300 python tools/stage_profile.py cfg2 2>&1 | cut -c1-100 > gpurun_out/r2c_stage_cfg2.log; cat gpurun_out/r2c_stage_cfg2.log
M2T_TIMING=1 python -m m2trans_b200.build --force > gpurun_out/az_build.log 2>&1
for cfg in cfg2 cfg1 cfg4; do timeout 300 python tools/az_timing.py $cfg > gpurun_out/az_timing_$cfg.log 2>&1; done
python -m m2trans_b200.build --force >> gpurun_out/az_build.log 2>&1
cat gpurun_out/az_timing_cfg2.log gpurun_out/az_timing_cfg1.log; tail -14 gpurun_out/az_timing_cfg4.log
