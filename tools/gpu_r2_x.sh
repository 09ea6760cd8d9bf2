#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_graph_gpu.py tests/test_stage_attnz_gpu.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2x_tests.log
for c in cfg1 cfg3 frame2; do timeout 300 python tools/stage_profile.py $c 2>&1 | grep "attn_z_kernelILi256\|replayed\|total" | cut -c1-100; done | tee gpurun_out/r2x_stage.log
timeout 300 python bench.py --workload cfg3 --no-cpu 2>&1 | tail -1 > gpurun_out/r2x_cfg3.log; python tools/show_bench.py < gpurun_out/r2x_cfg3.log
timeout 300 python bench.py --workload cfg1 --no-cpu 2>&1 | tail -1 > gpurun_out/r2x_cfg1.log; python tools/show_bench.py < gpurun_out/r2x_cfg1.log
