#!/bin/bash
# ncu --set full on one fused-MLP launch (stage 1, batch 32) of the MedCLIP pass
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_umma_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/clip_mlp -f python tools/bench_clip.py --steps 1 --warmup 1 > gpurun_out/ncu_clip_mlp.log 2>&1
tail -n 2 gpurun_out/ncu_clip_mlp.log
