#!/bin/bash
mkdir -p gpurun_out
M2T_TIMING=1 python -m m2trans_b200.build --force > gpurun_out/ct_build.log 2>&1
for c in ${CT_CFGS:-cfg2}; do timeout 300 python tools/conv_timing.py $c > gpurun_out/conv_timing_$c.log 2>&1; echo "== $c"; cat gpurun_out/conv_timing_$c.log; done
python -m m2trans_b200.build --force >> gpurun_out/ct_build.log 2>&1
