#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_graph_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2v_tests.log
timeout 300 python tools/stage_profile.py cfg2 2>&1 | grep "head_conv\|replayed\|total" | cut -c1-90
timeout 300 python tools/stage_profile.py cfg4 2>&1 | grep "head_conv\|replayed\|total" | cut -c1-90
timeout 300 python tools/stage_profile.py cfg1 2>&1 | grep "head_conv\|replayed\|total" | cut -c1-90
for i in 1 2; do timeout 300 python bench.py --no-cpu 2>&1 | tail -1 > gpurun_out/r2v_cfg2_$i.log; python tools/show_bench.py < gpurun_out/r2v_cfg2_$i.log; done
