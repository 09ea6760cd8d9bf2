#!/bin/bash
# usage: gpu_ncu_r2.sh <kernel-regex> <skip> <count> <outname> [workload]
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$1" -s $2 -c $3 -o gpurun_out/$4 -f python bench.py --steps 1 --warmup 3 --no-cpu --workload ${5:-cfg2} > gpurun_out/ncu_$4.log 2>&1
tail -2 gpurun_out/ncu_$4.log
ls -la gpurun_out/$4.ncu-rep
