#!/bin/bash
# clock64 phase stamps of attn_z: rebuild with M2T_TIMING=1 on the box, run, rebuild the product library
mkdir -p gpurun_out
M2T_TIMING=1 python -m m2trans_b200.build --force > gpurun_out/az_build.log 2>&1
for cfg in cfg2 cfg1 cfg4; do timeout 300 python tools/az_timing.py $cfg > gpurun_out/az_timing_$cfg.log 2>&1; done
python -m m2trans_b200.build --force >> gpurun_out/az_build.log 2>&1
cat gpurun_out/az_timing_cfg2.log gpurun_out/az_timing_cfg1.log gpurun_out/az_timing_cfg4.log
