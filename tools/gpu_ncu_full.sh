#!/bin/bash
# usage: gpu_ncu_full.sh <kernel-regex> <skip> <count> <outname>
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$1" -s $2 -c $3 -o gpurun_out/$4 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/$4.ncu-rep
