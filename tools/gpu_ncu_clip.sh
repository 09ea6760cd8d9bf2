#!/bin/bash
# ncu --set full on one stage-1 window-attention launch and one stage-1 fc1 (GELU) Linear of the MedCLIP pass (batch 32)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:clip_attn_kernel --launch-skip 12 --launch-count 1 -o gpurun_out/clip_attn -f python tools/bench_clip.py --steps 1 --warmup 1 > gpurun_out/ncu_clip_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lin_umma_kernel -c 1 --launch-skip 50 -o gpurun_out/clip_lin -f python tools/bench_clip.py --steps 1 --warmup 1 > gpurun_out/ncu_clip_lin.log 2>&1
tail -2 gpurun_out/ncu_clip_attn.log gpurun_out/ncu_clip_lin.log
