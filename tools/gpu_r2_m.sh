#!/bin/bash
# final evidence of the round: driver-style checks, then the ncu launch list + one --set full capture per kernel class
bash tools/gpu_final.sh
bash tools/gpu_r2_ncu.sh > gpurun_out/r2m_ncu.log 2>&1
tail -12 gpurun_out/r2m_ncu.log
