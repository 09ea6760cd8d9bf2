#!/bin/bash
# MedCLIP image pass: GPU tests, bench at batch 32 / 256, ncu launch list of one forward (batch 32)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_clip_gpu.py -q -s 2>&1 | grep -v "^\.*epi\|^$" | tail -25
timeout 120 python tools/bench_clip.py 2>&1 | tail -1
timeout 120 python tools/bench_clip.py --batch 256 --steps 5 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/clip_launches.csv python tools/bench_clip.py --steps 1 --warmup 0 > gpurun_out/clip_ncu.log 2>&1
tail -1 gpurun_out/clip_ncu.log
