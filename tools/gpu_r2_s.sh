#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tail_strip" -s 3 -c 1 -o gpurun_out/r02b_tail_strip -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02b_ncu_tail_strip.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tail_up_umma" -s 3 -c 1 -o gpurun_out/r02b_tail_up -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02b_ncu_tail_up.log 2>&1
ls -la gpurun_out/r02b*
