#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/stage_profile.py cfg4 2>&1 | cut -c1-110 | tail -10 | tee gpurun_out/r2q_stage_cfg4.log
timeout 300 python tools/stage_profile.py cfg3 2>&1 | cut -c1-110 | tail -10 | tee gpurun_out/r2q_stage_cfg3.log
