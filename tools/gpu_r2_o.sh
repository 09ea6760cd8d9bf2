#!/bin/bash
mkdir -p gpurun_out
M2T_TIMING=1 python -m m2trans_b200.build --force > gpurun_out/az_build.log 2>&1
timeout 300 python tools/az_timing.py cfg1 > gpurun_out/r2o_az_cfg1.log 2>&1
M2T_AZ_PAIRED=1 timeout 300 python tools/az_timing.py cfg1 > gpurun_out/r2o_az_cfg1_paired.log 2>&1
cat gpurun_out/r2o_az_cfg1.log gpurun_out/r2o_az_cfg1_paired.log
