#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_graph_gpu.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2ab_tests.log
for c in cfg3 cfg1 frame2; do timeout 300 python tools/stage_profile.py $c 2>&1 | grep "ffconv\|replayed\|total" | cut -c1-100; done | tee gpurun_out/r2ab_stage.log
for c in cfg3 cfg1; do timeout 300 python bench.py --workload $c --no-cpu 2>&1 | tail -1 > gpurun_out/r2ab_$c.log; python tools/show_bench.py < gpurun_out/r2ab_$c.log; done
