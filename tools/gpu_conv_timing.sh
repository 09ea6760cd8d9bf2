#!/bin/bash
mkdir -p gpurun_out
M2T_TIMING=1 python -m m2trans_b200.build --force > gpurun_out/ct_build.log 2>&1
timeout 300 python tools/conv_timing.py > gpurun_out/conv_timing.log 2>&1
python -m m2trans_b200.build --force >> gpurun_out/ct_build.log 2>&1
cat gpurun_out/conv_timing.log
